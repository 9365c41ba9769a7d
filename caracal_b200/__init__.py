"""caracal_b200 -- B200-native (sm_100a) RPMD hot path for Trebonius91/Caracal.

The product is the C-ABI shared library caracal_b200/libcaracal_gpu.so
(include/caracal_gpu.h).  This package only loads it (lib.py) and mirrors the reference's
operator interface for the path (api.py): egrad_<pes>(q,Natoms,Nbeads) -> V,dVdq,info ;
mdinit / verlet / recross / umbrella work units on batches of ring polymers.
"""
from .lib import (CaracalGpuError, LIB_PATH, PES_BRH2, PES_CH4H, PES_CH4OH, PES_GEH4OH, PES_CH4CN, PES_CLNH3, PES_NH3OH, PES_H2CO, PES_H3, PES_IDS, PES_O3, PES_OH3, TRANSFORM_EXACT,  # noqa: F401
                  TRANSFORM_REFERENCE, PATH_AUTO, PATH_FUSED, PATH_SPLIT, PES_HOSTCB, PES_QMDFF, PES_DGEVB, PES_WATER, PES_NONE, load)
from .api import (RPMD, Mechanism, UnimolMechanism, AtomShiftMechanism, atomic_mass_au, beta_calc_rate, beta_dynamic, dt_au, egrad, egrad_ch4h,  # noqa: F401
                  egrad_brh2, egrad_ch4cn, egrad_clnh3, egrad_nh3oh, egrad_h2co, egrad_ch4oh, egrad_geh4oh, egrad_h3, egrad_o3, egrad_oh3)


def build_if_needed(force=False):
    """Compile libcaracal_gpu.so in-tree if it is missing or stale (nvcc, sm_100a)."""
    from . import build as _b
    return _b.build(force=force)
