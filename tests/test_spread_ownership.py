"""The spread trajectory forms (csrc/traj_inst.cuh) hand every component of a bead to exactly one lane:
the ownership map and the per-lane evaluation of PesSpread compiled for the CPU (tests/host_harness/spread_host.cu).  No GPU
needed."""
import ctypes
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PES = {"h3": (1, 3), "oh3": (2, 4)}   # CRCL_PES_* id, atoms


@pytest.fixture(scope="module")
def spread(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("spread") / "libspread_host.so")
    src = os.path.join(ROOT, "tests", "host_harness", "spread_host.cu")
    subprocess.run(["nvcc", "-x", "cu", "-O1", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-DCRCL_FM_ON_HOST", "-w", "-o", so, src], check=True, capture_output=True)
    return ctypes.CDLL(so)


def test_ids_match_the_header():
    import re
    hdr = open(os.path.join(ROOT, "include", "caracal_gpu.h")).read()
    for name, (pid, _) in PES.items():
        assert int(re.search(r"#define CRCL_PES_%s (\d+)" % name.upper(), hdr).group(1)) == pid


@pytest.mark.parametrize("name", sorted(PES))
@pytest.mark.parametrize("lanes", [16, 8, 4, 2])
def test_every_component_has_exactly_one_owner(spread, name, lanes):
    pid, natoms = PES[name]
    nc = 3 * natoms
    nown = ctypes.c_int(0)
    seen = []
    for lane in range(lanes):
        for k in range(8):
            c = spread.hh_spread_owned(pid, lanes, lane, k, ctypes.byref(nown))
            assert c >= -1
            if k >= nown.value:
                assert c == -1          # nothing beyond the lane's slots
            if c >= 0:
                seen.append(c)
    assert nown.value == -(-nc // lanes)
    assert sorted(seen) == list(range(nc))
    # the dispatch of traj_inst.cuh pairs 16 / 8 / 4 / 2 lanes with 1 / 2 / 4 / 8 beads: 16 threads per trajectory, enough
    # for one pass of the cooperative xi (calc_xi_coop needs T >= 3 natoms)
    assert 16 >= nc


@pytest.mark.parametrize("name", sorted(PES))
@pytest.mark.parametrize("lanes", [16, 8, 4, 2])
def test_lanes_return_the_gradient_of_their_slots_and_lane_0_the_energy(spread, name, lanes):
    import numpy as np
    import sys
    sys.path.insert(0, ROOT)
    from caracal_b200 import systems
    pid, natoms = PES[name]
    nc = 3 * natoms
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(5)
    ts = np.asarray({"h3": systems.h3_ts, "oh3": systems.oh3_ts}[name](), dtype=np.float64).reshape(nc)
    for _ in range(5):
        q = np.ascontiguousarray(ts + 0.05 * rng.standard_normal(nc))
        v0, g0 = ctypes.c_double(0.0), np.zeros(nc)
        spread.hh_plain_eval(pid, q.ctypes.data_as(dp), ctypes.byref(v0), g0.ctypes.data_as(dp))
        got, esum, nown = np.full(nc, np.nan), 0.0, ctypes.c_int(0)
        for lane in range(lanes):
            v, gown = ctypes.c_double(0.0), np.zeros(8)
            spread.hh_spread_eval(pid, lanes, q.ctypes.data_as(dp), lane, ctypes.byref(v), gown.ctypes.data_as(dp))
            esum += v.value
            for k in range(8):
                c = spread.hh_spread_owned(pid, lanes, lane, k, ctypes.byref(nown))
                if c >= 0:
                    got[c] = gown[k]
        assert esum == v0.value            # lane 0 alone reports it
        assert np.array_equal(got, g0)     # the same arithmetic, selected
