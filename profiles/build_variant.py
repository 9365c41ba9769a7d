"""A/B builds: recompile ONE translation unit with extra flags and link it with the product's other objects into
caracal_b200/libcaracal_gpu_<suffix>.so (picked up through CRCL_LIB_PATH, see caracal_b200/lib.py).
    python profiles/build_variant.py nodmma traj_ch4h_recross.cu -DCRCL_DMMA_TRANSFORM=0"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caracal_b200 import build as B  # noqa: E402

suffix, tu, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
vdir = os.path.join(B.HERE, "build_" + suffix)
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, tu[:-3] + ".o")
subprocess.run(["nvcc"] + B.NVCC_FLAGS + flags + ["-c", "-o", obj, os.path.join(B.CSRC, tu)], check=True, capture_output=True)
objs = [obj if s == tu else os.path.join(B.OBJ, s[:-3] + ".o") for s in B._sources()]
lib = os.path.join(B.HERE, "libcaracal_gpu_%s.so" % suffix)
subprocess.run(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
print(lib)
