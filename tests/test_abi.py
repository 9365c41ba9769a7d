"""The C-ABI shared library loads and exports every symbol include/caracal_gpu.h declares; the
Python loader binds all of them; without a GPU the product fails loudly (no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "caracal_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crcl_[a-z0-9_]+)\s*\(", src)) - {"crcl_host_grad_fn"})


def test_header_symbols_exported():
    import caracal_b200
    caracal_b200.build_if_needed()
    out = subprocess.run(["nm", "-D", "--defined-only", caracal_b200.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (crcl_\w+)", out))
    declared = header_functions()
    assert len(declared) >= 20
    missing = [f for f in declared if f not in exported]
    assert not missing, "declared in caracal_gpu.h but not exported: %s" % missing
    extra = sorted(exported - set(declared))
    assert not extra, "exported but not declared in caracal_gpu.h: %s" % extra


def test_loader_binds_every_symbol():
    import caracal_b200
    from caracal_b200 import lib
    caracal_b200.build_if_needed()
    lib.load()
    assert sorted(lib.SIGNATURES) == header_functions()


def test_no_oracle_in_product():
    """The product path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "caracal_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
    out = subprocess.run(["ldd", os.path.join(pkg, "libcaracal_gpu.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import caracal_b200
    with pytest.raises(caracal_b200.CaracalGpuError, match="ENODEV"):
        caracal_b200.RPMD("h3", 16, [1837.0] * 3, 1000.0, 4.0)
    with pytest.raises(caracal_b200.CaracalGpuError):
        caracal_b200.egrad_h3([[0, 0, 0], [0, 0, 1.4], [0, 0, 4.0]])


def test_concurrent_builds_do_not_corrupt_the_library():
    """N ranks of a torchrun job all call build_if_needed(); with a stale translation unit they used to compile and link over
    each other and a rank would dlopen a half-written file (capture r2s: 'invalid ELF header' at N = 8).  Now a file lock
    serialises them and the link goes to a temporary name that is renamed into place: five processes, one stale unit."""
    import subprocess
    import sys
    from caracal_b200 import build as B
    B.build()
    stamp = os.path.join(B.OBJ, "water_kernels.o.sha")     # a unit that compiles in seconds
    assert os.path.exists(stamp)
    os.remove(stamp)
    code = "import caracal_b200, ctypes; caracal_b200.build_if_needed(); ctypes.CDLL(caracal_b200.LIB_PATH); print('ok')"
    procs = [subprocess.Popen([sys.executable, "-c", code], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for _ in range(5)]
    outs = [p.communicate(timeout=600) for p in procs]
    assert all(p.returncode == 0 and o[0].strip() == "ok" for p, o in zip(procs, outs)), [o[1][-300:] for o in outs]
    assert os.path.exists(stamp)
    assert not [f for f in os.listdir(B.HERE) if ".so.tmp." in f]
