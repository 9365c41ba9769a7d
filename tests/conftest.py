import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def host_harness():
    """The product's __host__ __device__ functors compiled for the CPU (tests/host_harness)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_harness")
    so = os.path.join(d, "libpes_host.so")
    src = os.path.join(d, "pes_host.cu")
    deps = [src] + [os.path.join(ROOT, "caracal_b200", "csrc", f)
                    for f in ("pes_h3.cuh", "pes_oh3.cuh", "pes_ch4h.cuh", "pes_brh2.cuh", "pes_o3.cuh", "pes_ch4oh.cuh", "pes_nh3x.cuh", "pes_h2co.cuh", "pes_h2co_tables.cuh", "xi.cuh", "rng.cuh",
                              "crcl_common.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.run(["nvcc", "-x", "cu", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                        "-Wno-deprecated-gpu-targets", "-DCRCL_FM_ON_HOST", "-o", so, src], check=True, capture_output=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def gpu():
    import caracal_b200
    caracal_b200.load()
    return caracal_b200
