"""Builds libcaracal_gpu.so in-tree with nvcc for sm_100a.

The shared library is the product: a C-ABI (include/caracal_gpu.h) over hand-written CUDA
kernels.  It is built in-tree (caracal_b200/libcaracal_gpu.so) so that it travels to the GPU
box with the repository snapshot.  Translation units are compiled in parallel and cached by
content hash of the unit and the headers it includes.
"""
import concurrent.futures
import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcaracal_gpu.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _includes(path, seen):
    if path in seen or not os.path.exists(path):
        return
    seen.add(path)
    for m in re.finditer(r'#include\s+"([^"]+)"', open(path).read()):
        _includes(os.path.normpath(os.path.join(os.path.dirname(path), m.group(1))), seen)


def _digest(src):
    seen = set()
    _includes(os.path.join(CSRC, src), seen)
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in sorted(seen):
        h.update(open(p, "rb").read())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(src)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, "", False
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, os.path.join(CSRC, src)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = " ".join(cmd) + "\n" + res.stdout + res.stderr
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s" % (src, log[-6000:]))
    with open(stamp, "w") as f:
        f.write(dig)
    return obj, log, True


def build(force=False, verbose=False):
    """Safe to call from every rank of a multi-process job at once: an exclusive file lock serialises the callers (the
    first one compiles, the others find everything up to date), and the library is linked to a temporary name and
    renamed into place, so a process that is loading it never sees a half-written file."""
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(force, verbose):
    srcs = _sources()
    if force:
        for s in srcs:
            st = os.path.join(OBJ, s[:-3] + ".o.sha")
            if os.path.exists(st):
                os.remove(st)
    logs, rebuilt = [], False
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        for obj, log, did in ex.map(_compile, srcs):
            logs.append(log)
            rebuilt = rebuilt or did
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if rebuilt or not os.path.exists(LIB):
        nvcc = os.environ.get("NVCC", "nvcc")
        tmp = LIB + ".tmp.%d" % os.getpid()
        # no library besides the CUDA runtime: the FFT of ewald_recip.f90 is a kernel of this library too
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
        os.replace(tmp, LIB)
        with open(os.path.join(HERE, "build.log"), "a" if not force else "w") as f:
            f.write("\n".join(l for l in logs if l))
    if verbose:
        print("rebuilt" if rebuilt else "up to date", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
