"""Kernel span of the BASE step kernel per CTA (experimental build): entry / first env / exit times relative to the earliest entry."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
args = bench.parse()
import torch
from predpreygrass_b200 import _lib
from predpreygrass_b200.batched import BatchedPredPreyGrass
L = _lib.load()
fn = L.ppg_debug_span_base; fn.argtypes = [C.c_void_p]; fn.restype = C.c_int
env = BatchedPredPreyGrass(bench.build_config(args, seed=1000), args.envs, device=0)
env.reset()
for _ in range(args.preroll):
    a0, a1 = env.random_actions(4242); env.step(a0, a1)
torch.cuda.synchronize()
for rep in range(3):
    a0, a1 = env.random_actions(4242)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.step(a0, a1); e1.record(); torch.cuda.synchronize()
    buf = np.zeros((3, 8192), np.uint32); assert fn(buf.ctypes.data) == 0
    n = min(8192, 148 * 32)
    ent, first, ex = [b[b > 0].astype(np.int64) for b in buf]
    t0 = ent.min()
    print(f"step+obs by events {e0.elapsed_time(e1) * 1e3:.1f} us | CTAs {len(ent)}: entry {np.percentile(ent - t0, [0, 50, 100]) / 1e3} us, "
          f"first env start {np.percentile(first - t0, [0, 50, 100]) / 1e3} us, exit {np.percentile(ex - t0, [0, 50, 99, 100]) / 1e3} us")
