"""Dynamic operation census of the QMDFF / DG-EVB restatement (BASELINE.md section 4 convention: add, mul, div, sqrt = 1,
every libm call = 1; pair and donor-acceptor tests that fail their cut-off are not counted), from the
counting build liboracle_count.so (count_qmdff.cpp).  Writes oracle/flop_census_qmdff.json.  TEST INFRASTRUCTURE.
    python oracle/census_qmdff.py"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from tests.qmdff_synth import HEXANE, make_dgevb, make_system  # noqa: E402
from oracle import oracle as O  # noqa: E402

so = os.path.join(HERE, "liboracle_count.so")
subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", "-o", so, os.path.join(HERE, "count_qmdff.cpp")], check=True, cwd=HERE)
L = ctypes.CDLL(so)


def counted(fn):
    L.oracle_count_reset()
    fn()
    out = (ctypes.c_ulonglong * 6)()
    L.oracle_count_get(out)
    add, mul, div, sq, libm, cmp_ = [int(v) for v in out]
    return dict(add=add, mul=mul, div=div, sqrt=sq, libm=libm, flops=add + mul + div + sq + libm)


def qmdff_flops(T, x):
    Q = O.Qmdff(T)
    Q.L = L
    L.orc_qmdff_egrad.argtypes = [ctypes.POINTER(O._QmdffStruct), O.dp, ctypes.c_int, O.dp, O.dp]
    return counted(lambda: Q.egrad(x))


def only(T, keep):
    """T with every list but `keep` emptied (single molecule, non-periodic: list terms only)"""
    T = dict(T)
    for k, w, v, wv in (("bond", 2, "vbond", 3), ("angl", 3, "vangl", 2), ("tors", 6, "vtors", T["ldvt"]), ("nci", 3, None, 0)):
        if k != keep:
            T[k] = np.zeros((0, w), dtype=np.int32)
            if v:
                T[v] = np.zeros((0, wv))
    return T


out = {"convention": "BASELINE.md section 4: add/mul/div/sqrt/libm = 1 each; rejected cut-off tests not counted"}
# list terms: one ethanol-like molecule, gas phase
T1 = make_system(nmol=1, seed=3, periodic=False, zahn=False, hb=False, frac_formaldehyde=0.0)
x1 = T1["xyz"][None] + np.random.default_rng(0).normal(0, 0.05, (4,) + T1["xyz"].shape)
for key, name in (("bond", "bond"), ("angl", "angle"), ("tors", "torsion"), ("nci", "nci_pair")):
    Tk = only(T1, key)
    n = len(Tk[key])
    if key == "nci" and n <= 1:
        continue
    c = qmdff_flops(Tk, x1)
    out[name] = dict(per_term=c["flops"] / (4.0 * n), libm_per_term=c["libm"] / (4.0 * n), terms=n)
# inter-molecular pair, periodic Zahn: two molecules of one atom each, inside / outside the cut-offs
T2 = make_system(nmol=2, seed=3, periodic=True, zahn=True, hb=False, frac_formaldehyde=0.0)
na = T2["n"] // 2
sel = [0, na]                                             # first atom of each molecule
Tp = only(T2, None)
Tp.update(n=2, at=np.asarray(T2["at"])[sel], q=np.asarray(T2["q"])[sel], molnum=np.array([1, 2], dtype=np.int32), nmols=2,
          c6xy=np.asfortranarray(np.asarray(T2["c6xy"])[np.ix_(sel, sel)]))
if "q_glob" in Tp:
    Tp["q_glob"] = np.asarray(Tp["q_glob"])[sel]
xin = np.array([[[1.0, 1.0, 1.0], [6.0, 2.0, 1.5]]])
c_in = qmdff_flops(Tp, xin)
xout = np.array([[[1.0, 1.0, 1.0], [1.0 + 0.45 * T2["box"][0], 1.0 + 0.45 * T2["box"][1], 1.0 + 0.45 * T2["box"][2]]]])
c_out = qmdff_flops(Tp, xout)
out["inter_pair_vdw_plus_zahn"] = dict(per_pair=c_in["flops"], libm_per_pair=c_in["libm"], rejected_pair=c_out["flops"])
# the config-5 bench system: whole image, per atom (lists + pairs inside the cut-offs + H-bond terms)
for hb in (False, True):
    Tb = make_system(nmol=385, seed=12, periodic=True, zahn=True, hb=hb)
    xb = Tb["xyz"][None] + np.random.default_rng(3).normal(0, 0.05, (1,) + Tb["xyz"].shape)
    c = qmdff_flops(Tb, xb)
    out["box_385_molecules_hb%d" % int(hb)] = dict(natoms=int(Tb["n"]), flops_per_image=c["flops"], libm_per_image=c["libm"],
                                                   flops_per_atom=c["flops"] / float(Tb["n"]))
# config 4: the 20-atom two-state DG-EVB system (mode 3, 7 Gaussians, nat6 = 12), one image
T1h, T2h, E = make_dgevb(seed=1, mode=3, npoints=7, template=HEXANE)
D = O.Dgevb(T1h, T2h, E)
D.L = L
for name in dir(O.lib()):
    pass
xh = T1h["xyz"][None] + np.random.default_rng(5).normal(0, 0.03, (4,) + T1h["xyz"].shape)
try:
    L.orc_dgevb_egrad.argtypes = O.lib().orc_dgevb_egrad.argtypes
    c = counted(lambda: D.egrad(xh))
    out["dgevb_hexane_mode3_7points"] = dict(natoms=int(T1h["n"]), flops_per_image=c["flops"] / 4.0, libm_per_image=c["libm"] / 4.0)
except Exception as exc:                                   # pragma: no cover
    out["dgevb_hexane_mode3_7points"] = dict(error=str(exc))
json.dump(out, open(os.path.join(HERE, "flop_census_qmdff.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
