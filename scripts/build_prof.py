"""Experimental build with the per-phase latency profile of the step kernels compiled in (-DPPG_PHASE_PROF):
predpreygrass_b200/libppg_b200_prof.so, loaded with PPG_LIB=... (scripts/phase_profile.py).  Not the shipped library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from predpreygrass_b200 import build as B

OUT = os.path.join(B.HERE, "libppg_b200_prof.so")
extra = ["-DPPG_PHASE_PROF"] + sys.argv[1:]


def one(src):
    obj = os.path.join(B.CSRC, src[:-3] + ".prof.o")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", obj])
    return obj


with ThreadPoolExecutor(max_workers=5) as ex:
    objs = list(ex.map(one, B.SOURCES))
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
print(OUT)
