"""Time the attention kernels at the cfg-2 shapes (CUDA events, L2 flushed between launches)."""
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


for (name, B, T, H, d, drop) in (("teacher fwd64", 32, 779, 12, 64, None), ("student fwd40", 32, 389, 12, 40, None),
                                  ("student fwd40 drop", 32, 389, 12, 40, (123, 0.1))):
    qkv = torch.randn(B, T, 3 * H * d, device=dev).to(h)
    vt = torch.tensor([T - 3 * i for i in range(B)], device=dev, dtype=torch.int32)
    out, lse = torch.empty(B * T, H * d, device=dev, dtype=h), torch.empty(B, H, T, device=dev)
    us = timeit(lambda: K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5, drop=drop))
    fl = 4.0 * sum(int(v) ** 2 for v in vt.tolist()) * d * H
    print(f"{name}: {us:.1f} us  {fl / us / 1e6:.0f} TF/s (valid keys only)")
    if "fwd40" in name:
        do, dqkv, delta = torch.randn(B, T, H * d, device=dev).to(h), torch.empty_like(qkv), torch.empty(B, H, T, device=dev)
        ws = torch.empty(B * T, H * d, device=dev)
        us = timeit(lambda: K.attn_bwd(qkv, vt, out, do, lse, dqkv, delta, B, T, H, d, d ** -0.5, dq_ws=ws, drop=drop))
        print(f"  bwd40{' drop' if drop else ''}: {us:.1f} us (delta + fused bwd + dq convert)")
