"""The synthetic QMDFF generators live in the package (caracal_b200/qmdff_synth.py: bench.py uses them too); the tests
keep importing them from here.  make_dgevb places its Gaussian centres with the CPU restatement's xyz_2int, as the
fixtures were made."""
import numpy as np

from caracal_b200 import qmdff_synth as _S
from caracal_b200.qmdff_synth import *  # noqa: F401,F403
from caracal_b200.qmdff_synth import HEXANE, make_system  # noqa: F401


def make_dgevb(seed=0, mode=3, npoints=5, template=None):
    from oracle import oracle as O
    box = {}

    def internals(coord_def, xyz):
        if "d" not in box:
            T1 = _S.make_system(nmol=1, seed=seed, periodic=False, frac_formaldehyde=0.0, hb=True, template=template)
            nat6 = len(coord_def)
            box["d"] = O.Dgevb(T1, T1, dict(mode=mode, coord_def=coord_def, g_thres=1e-10, point_int=np.zeros((1, nat6)),
                                            alph=np.ones(1), b_vec=np.zeros(4000)))
        return box["d"].internals(xyz)
    return _S.make_dgevb(seed, mode, npoints, template, internals=internals)
