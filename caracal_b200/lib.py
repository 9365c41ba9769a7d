"""ctypes loader for libcaracal_gpu.so (the C-ABI in include/caracal_gpu.h).

There is no fallback: if the library is missing or no CUDA device is present the calls fail
loudly.  Nothing under oracle/ is ever imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CRCL_LIB_PATH selects an alternative build of the same library (kernel-tuning experiments)
LIB_PATH = os.environ.get("CRCL_LIB_PATH", os.path.join(_HERE, "libcaracal_gpu.so"))

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_u32_p = ctypes.POINTER(ctypes.c_uint32)

PES_NONE, PES_H3, PES_OH3, PES_CH4H, PES_BRH2, PES_O3, PES_QMDFF, PES_DGEVB, PES_WATER, PES_HOSTCB = 0, 1, 2, 3, 4, 5, 10, 11, 12, 100
PES_CH4OH, PES_GEH4OH, PES_CH4CN, PES_CLNH3, PES_NH3OH, PES_H2CO = 6, 7, 8, 9, 13, 14
PES_IDS = {"h3": PES_H3, "oh3": PES_OH3, "ch4h": PES_CH4H, "brh2": PES_BRH2, "o3": PES_O3, "ch4oh": PES_CH4OH, "geh4oh": PES_GEH4OH, "ch4cn": PES_CH4CN, "clnh3": PES_CLNH3, "nh3oh": PES_NH3OH, "h2co": PES_H2CO}
PES_NATOMS = {PES_H3: 3, PES_OH3: 4, PES_CH4H: 6, PES_BRH2: 3, PES_O3: 3, PES_CH4OH: 7, PES_GEH4OH: 7, PES_CH4CN: 7, PES_CLNH3: 5, PES_NH3OH: 6, PES_H2CO: 4}
TRANSFORM_REFERENCE, TRANSFORM_EXACT = 0, 1
PATH_AUTO, PATH_FUSED, PATH_SPLIT = 0, 1, 2
ERRORS = {0: "CRCL_OK", -1: "CRCL_ENODEV", -2: "CRCL_EINVAL", -3: "CRCL_ENOMEM", -4: "CRCL_ECUDA",
          -5: "CRCL_ENOSUP", -6: "CRCL_ESTATE"}
TRAJ_OK, TRAJ_SHAKE_FAIL, TRAJ_NAN, TRAJ_SINGULAR, TRAJ_ENERGY, TRAJ_PESWARN, TRAJ_XI_RANGE, TRAJ_PBC_FAIL = 0, 1, 2, 4, 8, 16, 32, 64
TRAJ_FATAL = 1 | 2 | 4 | 8 | 32 | 64

class WaterParams(ctypes.Structure):
    """crcl_water_params of include/caracal_gpu.h"""
    _fields_ = [("n", ctypes.c_int), ("periodic", ctypes.c_int), ("zahn", ctypes.c_int), ("box", ctypes.c_double * 3),
                ("coul_cut", ctypes.c_double), ("zahn_a", ctypes.c_double), ("zahn_par", ctypes.c_double),
                ("pars", ctypes.c_double * 11), ("q", c_double_p), ("is_O", c_int_p)]


class QmdffTables(ctypes.Structure):
    """crcl_qmdff_tables of include/caracal_gpu.h"""
    _fields_ = [("n", ctypes.c_int), ("at", c_int_p), ("q", c_double_p), ("molnum", c_int_p), ("nmols", ctypes.c_int),
                ("nbond", ctypes.c_int), ("nangl", ctypes.c_int), ("ntors", ctypes.c_int), ("nhb", ctypes.c_int),
                ("nnci", ctypes.c_int), ("ldvt", ctypes.c_int),
                ("bond", c_int_p), ("vbond", c_double_p), ("angl", c_int_p), ("vangl", c_double_p),
                ("tors", c_int_p), ("vtors", c_double_p), ("nci", c_int_p), ("c6xy", c_double_p),
                ("r0ab", c_double_p), ("zab", c_double_p), ("r094", c_double_p), ("sr42", c_double_p),
                ("rad", c_double_p), ("eps1", ctypes.c_double * 6), ("eps2", ctypes.c_double * 6),
                ("periodic", ctypes.c_int), ("zahn", ctypes.c_int), ("box", ctypes.c_double * 3),
                ("coul_cut", ctypes.c_double), ("vdw_cut", ctypes.c_double), ("cut_low", ctypes.c_double),
                ("zahn_a", ctypes.c_double), ("zahn_par", ctypes.c_double), ("e_zero", ctypes.c_double),
                ("hb", c_int_p), ("vhb", c_double_p), ("scalehb", c_double_p), ("scalexb", c_double_p),
                ("q_glob", c_double_p)]


class DgevbParams(ctypes.Structure):
    """crcl_dgevb_params of include/caracal_gpu.h"""
    _fields_ = [("mode", ctypes.c_int), ("npoints", ctypes.c_int), ("nat6", ctypes.c_int), ("coord_def", c_int_p),
                ("point_int", c_double_p), ("alph", c_double_p), ("b_vec", c_double_p), ("g_thres", ctypes.c_double)]


class EwaldParams(ctypes.Structure):
    """crcl_ewald_params of include/caracal_gpu.h"""
    _fields_ = [("box", ctypes.c_double * 3), ("a_ewald", ctypes.c_double), ("nfft", ctypes.c_int),
                ("bsorder", ctypes.c_int), ("bsmod1", c_double_p), ("bsmod2", c_double_p), ("bsmod3", c_double_p)]


# every symbol include/caracal_gpu.h declares: (restype, argtypes)
_H = ctypes.c_void_p
SIGNATURES = {
    "crcl_create": (ctypes.c_int, [ctypes.POINTER(_H), ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p,
                                   c_int_p, ctypes.c_double, ctypes.c_double, ctypes.c_int]),
    "crcl_destroy": (ctypes.c_int, [_H]),
    "crcl_last_error": (ctypes.c_char_p, [_H]),
    "crcl_set_stream": (ctypes.c_int, [_H, ctypes.c_void_p]),
    "crcl_synchronize": (ctypes.c_int, [_H]),
    "crcl_set_beta_dt": (ctypes.c_int, [_H, ctypes.c_double, ctypes.c_double]),
    "crcl_set_transform": (ctypes.c_int, [_H, ctypes.c_int]),
    "crcl_set_spread_max_beads": (ctypes.c_int, [_H, ctypes.c_int]),
    "crcl_set_host_gradient_cb": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_void_p]),
    "crcl_set_qmdff": (ctypes.c_int, [_H, ctypes.POINTER(QmdffTables)]),
    "crcl_set_qmdff2": (ctypes.c_int, [_H, ctypes.POINTER(QmdffTables)]),
    "crcl_set_dgevb": (ctypes.c_int, [_H, ctypes.POINTER(DgevbParams)]),
    "crcl_set_path": (ctypes.c_int, [_H, ctypes.c_int]),
    "crcl_set_graph": (ctypes.c_int, [_H, ctypes.c_int]),
    "crcl_set_water": (ctypes.c_int, [_H, ctypes.POINTER(WaterParams)]),
    "crcl_set_ewald": (ctypes.c_int, [_H, ctypes.POINTER(EwaldParams)]),
    "crcl_ewald_recip": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                        c_double_p]),
    "crcl_set_mechanism": (ctypes.c_int, [_H, ctypes.c_int, c_int_p, ctypes.c_int, c_int_p, c_double_p, c_double_p,
                                          ctypes.c_int, c_int_p, c_int_p, ctypes.c_double]),
    "crcl_set_mechanism_unimol": (ctypes.c_int, [_H, ctypes.c_int, c_int_p, ctypes.c_int, c_int_p, c_double_p, c_double_p,
                                                 c_double_p, c_double_p]),
    "crcl_set_mechanism_atom_shift": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                                     ctypes.c_double, ctypes.c_double]),
    "crcl_set_thermostat": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double]),
    "crcl_set_seed": (ctypes.c_int, [_H, ctypes.c_uint64]),
    "crcl_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "crcl_comm_init": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "crcl_comm_destroy": (ctypes.c_int, [_H]),
    "crcl_comm_info": (ctypes.c_int, [_H, c_int_p, c_int_p, c_int_p]),
    "crcl_set_box": (ctypes.c_int, [_H, ctypes.c_int, c_double_p]),
    "crcl_set_rpmd_check": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double]),
    "crcl_egrad": (ctypes.c_int, [_H, ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                  c_int_p]),
    "crcl_egrad_dev": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_void_p]),
    "crcl_verlet": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                   c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_int_p,
                                   c_u32_p, c_u32_p]),
    "crcl_verlet_dev": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 11),
    "crcl_mdinit": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p,
                                   c_double_p, c_double_p, c_u32_p, c_u32_p]),
    "crcl_calc_xi": (ctypes.c_int, [_H, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                                    c_double_p]),
    "crcl_recross_children": (ctypes.c_int, [_H, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, c_double_p, c_double_p, c_int_p]),
    "crcl_recross_children_dev": (ctypes.c_int, [_H, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
                                                 ctypes.c_void_p]),
    "crcl_umbrella_window": (ctypes.c_int, [_H, c_double_p, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_uint32, c_double_p, c_double_p,
                                            c_int_p]),
    "crcl_umbrella_windows": (ctypes.c_int, [_H, ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, c_double_p,
                                             c_double_p, c_int_p]),
    "crcl_rng_normals": (ctypes.c_int, [_H, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                        ctypes.c_int, c_double_p]),
    "crcl_launch_count": (ctypes.c_longlong, [_H]),
    "crcl_last_kernel_ms": (ctypes.c_double, [_H]),
    "crcl_kernel_timings": (ctypes.c_int, [_H, c_double_p, ctypes.c_int]),
    "crcl_bench_propagate": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, c_double_p]),
    "crcl_measure_fp64_tflops": (ctypes.c_double, [_H, ctypes.c_int]),
    "crcl_measure_dmma_tflops": (ctypes.c_double, [_H, ctypes.c_int]),
    "crcl_bench_transform": (ctypes.c_int, [_H, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p]),
}



_lib = None


class CaracalGpuError(RuntimeError):
    pass


def _prefer_bundled_nccl():
    """The library binds NCCL at run time (dlopen of libnccl.so.2, csrc/comm.cuh).  In a Python process that also
    imports PyTorch the two must agree on ONE libnccl.so.2: whichever is loaded first wins for the whole process (the
    loader deduplicates by soname), and PyTorch does not import against an older system NCCL.  So point the library
    at the pip-installed copy PyTorch itself uses, unless the user chose one (CRCL_NCCL_LIB).  No torch import here."""
    if os.environ.get("CRCL_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CRCL_NCCL_LIB"] = cand
                return
    except (ImportError, ValueError, AttributeError):
        pass


def load():
    """Load libcaracal_gpu.so and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CaracalGpuError(
            "%s not found: build it with `python -m caracal_b200.build` (no CPU fallback exists)" % LIB_PATH)
    _prefer_bundled_nccl()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, handle=None, what=""):
    if rc == 0:
        return
    msg = ""
    if handle:
        msg = (load().crcl_last_error(handle) or b"").decode()
    raise CaracalGpuError("%s failed: %s %s" % (what, ERRORS.get(rc, str(rc)), msg))
