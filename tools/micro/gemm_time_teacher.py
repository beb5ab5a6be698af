"""Time the teacher's 24 928-row GEMMs (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fithubert_b200 import kernels as K

dev, h = "cuda", torch.float16
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


M, E, F = 32 * 779, 768, 3072
x = torch.randn(M, E, device=dev).to(h); xf = torch.randn(M, F, device=dev).to(h)
res = torch.randn(M, E, device=dev).to(h)
wq = (0.03 * torch.randn(3 * E, E, device=dev)).to(h); wo = (0.03 * torch.randn(E, E, device=dev)).to(h)
w1 = (0.03 * torch.randn(F, E, device=dev)).to(h); w2 = (0.03 * torch.randn(E, F, device=dev)).to(h)
bq, bo, b1 = torch.randn(3 * E, device=dev), torch.randn(E, device=dev), torch.randn(F, device=dev)
oq, oo, o1 = torch.empty(M, 3 * E, device=dev, dtype=h), torch.empty(M, E, device=dev, dtype=h), torch.empty(M, F, device=dev, dtype=h)
for name, fn, n, k in (("qkv 768 -> 2304", lambda: K.linear(x, wq, bq, out=oq), 3 * E, E),
                       ("out_proj 768 -> 768 + residual", lambda: K.linear(x, wo, bo, out=oo, residual=res), E, E),
                       ("fc1 768 -> 3072 gelu", lambda: K.linear(x, w1, b1, out=o1, gelu=True), F, E),
                       ("fc2 3072 -> 768 + residual", lambda: K.linear(xf, w2, bo, out=oo, residual=res), E, F)):
    us = timeit(fn)
    print(f"{name:34s} {us:6.1f} us  {2.0 * M * n * k / us / 1e6:6.0f} TF/s")
