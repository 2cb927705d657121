"""write-only HBM bandwidth of this GPU (fill / memset over 2 GiB) next to the read+write copy of MEASURED_PEAKS.json —
the observation kernel is a write stream, so this is the ceiling its `achieved` GB/s can actually approach"""
import torch
x = torch.empty(1 << 29, dtype=torch.float32, device="cuda")  # 2 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: x.fill_(1.0)); print("fill_  2 GiB: %.3f ms  %.0f GB/s (write only)" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: x.zero_()); print("zero_  2 GiB: %.3f ms  %.0f GB/s (write only)" % (ms, x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.copy_(x)); print("copy_  2 GiB: %.3f ms  %.0f GB/s (read + write)" % (ms, 2 * x.numel() * 4 / ms / 1e6))
ms = t(lambda: x.sum()); print("sum    2 GiB: %.3f ms  %.0f GB/s (read only)" % (ms, x.numel() * 4 / ms / 1e6))
