"""Time the student's 12 448-row GEMM variants (CUDA events, L2 flushed between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from fithubert_b200 import kernels as K

dev, h, f = "cuda", torch.float16, torch.float32
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=30):
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


M, E = 32 * 389, 480
x = (torch.randn(M, E, device=dev)).to(h)
w = (0.05 * torch.randn(E, E, device=dev)).to(h)
w3 = (0.05 * torch.randn(3 * E, E, device=dev)).to(h)
b, b3 = torch.randn(E, device=dev), torch.randn(3 * E, device=dev)
res32, out32 = torch.randn(M, E, device=dev), torch.empty(M, E, device=dev)
out16, u16 = torch.empty(M, E, device=dev, dtype=h), torch.empty(M, E, device=dev, dtype=h)
qkv = torch.empty(M, 3 * E, device=dev, dtype=h)
dy = torch.randn(M, E, device=dev).to(h)
dqkv = torch.randn(M, 3 * E, device=dev).to(h)
cases = [
    ("fwd qkv      480 -> 1440, bias", lambda: K.linear(x, w3, b3, out=qkv), 1440),
    ("fwd out_proj + fp32 residual -> fp32", lambda: K.linear(x, w, b, out=out32, residual=res32, out_dtype=f), 480),
    ("fwd out_proj + fp32 residual + dropout", lambda: K.linear(x, w, b, out=out32, residual=res32, out_dtype=f, drop=(5, 0.1)), 480),
    ("fwd fc1 gelu + saved gelu'", lambda: K.linear(x, w, b, gelu=True, dgelu_out=u16, out=out16), 480),
    ("bwd dgrad plain", lambda: K.linear_dgrad(dy, w, out=out16), 480),
    ("bwd dgrad x saved gelu'", lambda: K.linear_dgrad(dy, w, mul_aux=u16, out=out16), 480),
    ("bwd dgrad + fp32 residual -> fp32", lambda: K.linear_dgrad(dy, w, residual=res32, out=out32, out_dtype=f), 480),
    ("bwd qkv dgrad 1440 -> 480 + fp32 residual -> fp32", lambda: K.linear_dgrad(dqkv, w3, residual=res32, out=out32, out_dtype=f), 1440),
]
for name, fn, n in cases:
    us = timeit(fn)
    print(f"{name:52s} {us:6.1f} us  {2.0 * M * E * n / us / 1e6:6.0f} TF/s")
