"""Experimental build of the library with extra nvcc flags: python scripts/build_variant.py NAME -DPPG_MIN_CTAS=24 ...
-> predpreygrass_b200/libppg_b200_NAME.so (load with PPG_LIB=...).  Not the shipped library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from predpreygrass_b200 import build as B

name, extra = sys.argv[1], sys.argv[2:]
OUT = os.path.join(B.HERE, f"libppg_b200_{name}.so")


def one(src):
    obj = os.path.join(B.CSRC, src[:-3] + f".{name}.o")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc"] + B.NVCC_FLAGS + extra + ["-c", os.path.join(B.CSRC, src), "-o", obj])
    return obj


with ThreadPoolExecutor(max_workers=5) as ex:
    objs = list(ex.map(one, B.SOURCES))
subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
for o in objs:
    os.remove(o)
print(OUT)
