"""Time the LayerNorm / column-sum kernels at the cfg-2 shapes against a plain copy of the same bytes
(CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fithubert_b200 import kernels as K

dev, h, f = "cuda", torch.float16, torch.float32
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=30, cold=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if cold:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def report(name, nbytes, fn):
    for cold in (True, False):
        us = timeit(fn, cold=cold)
        print(f"{name:46s} {'cold' if cold else 'warm'} {us:7.1f} us  {nbytes / us / 1e3:7.0f} GB/s")


# teacher LayerNorm fp16 -> fp16
rows, C = 24928, 768
x = torch.randn(rows, C, device=dev).to(h); y = torch.empty_like(x)
g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
report("layernorm_fwd teacher 24928x768 fp16", 2 * x.numel() * 2, lambda: K.layernorm_fwd(x, g, b, y))
report("  copy of the same bytes", 2 * x.numel() * 2, lambda: y.copy_(x))
# student LayerNorm fp32 in, fp16 + fp32 out (+ stats)
rows, C = 12448, 480
x32 = torch.randn(rows, C, device=dev); y16 = torch.empty(rows, C, device=dev, dtype=h); y32 = torch.empty_like(x32)
g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
report("layernorm_fwd32 student 12448x480", x32.numel() * 10, lambda: K.layernorm_fwd32(x32, g, b, y16, y32, mean, rstd))
report("  copy fp32 + fp16 of the same bytes", x32.numel() * 10, lambda: (y32.copy_(x32), y16.copy_(y16)))
# student LayerNorm backward: dy32 + dy2 (fp16) + x32 in, dx fp16 + dx32 out
dy32 = torch.randn(rows, C, device=dev); dy2 = torch.randn(rows, C, device=dev).to(h)
K.layernorm_fwd32(x32, g, b, y16, y32, mean, rstd)
dx = torch.empty(rows, C, device=dev, dtype=h); dx32 = torch.empty_like(x32)
dg, db, ds = (torch.zeros(C, device=dev) for _ in range(3))
report("layernorm_bwd32 student", x32.numel() * 16, lambda: K.layernorm_bwd32(dy32, x32, g, mean, rstd, dg, db, dy2=dy2, dx=dx, dx32=dx32, dxsum=ds))
report("layernorm_bwd32 student, no dy2", x32.numel() * 14, lambda: K.layernorm_bwd32(dy32, x32, g, mean, rstd, dg, db, dx=dx, dx32=dx32, dxsum=ds))
# column sums
t = torch.randn(rows, 1440, device=dev).to(h); out = torch.zeros(1440, device=dev)
report("colsum 12448x1440", t.numel() * 2, lambda: K.colsum(t, out))
t2 = torch.randn(rows, 480, device=dev).to(h); out2 = torch.zeros(480, device=dev)
report("colsum 12448x480", t2.numel() * 2, lambda: K.colsum(t2, out2))
