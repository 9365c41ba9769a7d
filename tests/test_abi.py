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
