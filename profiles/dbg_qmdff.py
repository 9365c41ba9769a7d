import numpy as np, sys
sys.path.insert(0,'.')
import caracal_b200 as gpu
from oracle import oracle as O
from tests import common as C
from tests.qmdff_synth import make_system
from tests.test_gpu_qmdff import handle
T = make_system(nmol=8, seed=9, periodic=True, zahn=False)
rng = np.random.default_rng(1)
x = T["xyz"][None] + rng.normal(0, 0.06, (48,) + T["xyz"].shape)
def cmp(T, label):
    g,_ = handle(gpu, T); Q = O.Qmdff(T)
    Vo, go = Q.egrad(x); Vd, gd, _ = g.egrad(x); gd = gd.reshape(go.shape)
    err = np.abs(gd-go).max(axis=(1,2)); i = err.argmax()
    a = np.abs(gd[i]-go[i]).max(axis=1).argmax()
    print(label, 'max abs err', err.max(), 'image', i, 'atom', a, 'Z', T['at'][a], 'mol', T['molnum'][a], 'gmax', np.abs(go[i]).max())
    return i, a
cmp(T, 'full')
Tb = dict(T); Tb['nci'] = T['nci'][:0]; Tb['nmols'] = 1; cmp(Tb, 'bonded only')
Tn = dict(T); Tn['bond']=T['bond'][:0]; Tn['vbond']=T['vbond'][:0]; Tn['angl']=T['angl'][:0]; Tn['vangl']=T['vangl'][:0]; Tn['tors']=T['tors'][:0]; Tn['vtors']=T['vtors'][:0]
cmp(Tn, 'nonbonded only')
Tn2 = dict(Tn); Tn2['nmols']=1; cmp(Tn2, 'nci only')
Tn3 = dict(Tn); Tn3['q'] = T['q']*0; cmp(Tn3, 'nonbonded, no charges')
for sel,name in [(lambda t: t[5]!=2,'proper torsions only'),(lambda t: t[5]==2,'inversions only')]:
    Tt = dict(Tb); m = np.array([sel(t) for t in T['tors']]); Tt['tors']=T['tors'][m]; Tt['vtors']=T['vtors'][m]
    Tt['bond']=T['bond'][:0]; Tt['vbond']=T['vbond'][:0]; Tt['angl']=T['angl'][:0]; Tt['vangl']=T['vangl'][:0]
    cmp(Tt, name)
Ta = dict(Tb); Ta['tors']=T['tors'][:0]; Ta['vtors']=T['vtors'][:0]; Ta['bond']=T['bond'][:0]; Ta['vbond']=T['vbond'][:0]; cmp(Ta,'angles only')
