"""Isolated timings of the attention-map kernels at the ex.yaml / cfg-2 size (B = 32, H = 12, T = 779, d = 64).
usage: python tools/attn_kernel_bench.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K

B, H, T, d = 32, 12, 779, 64
E = H * d
dev = "cuda"
qkv = (torch.randn(B * T, 3 * E, device=dev) * 0.5).half()
tqkv = (torch.randn(B * T, 3 * E, device=dev) * 0.5).half()
valid = torch.tensor([T - 3 * i for i in range(B)], dtype=torch.int32, device=dev)
q, k, v = qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:]
S = K.attn_scores(q, k, valid, B, T, H, d, d ** -0.5)
St = K.attn_scores(tqkv[:, :E], tqkv[:, E:2 * E], valid, B, T, H, d, d ** -0.5)
G = torch.empty(S.shape, device=dev, dtype=torch.float16)
loss = torch.zeros(1, device=dev)
dqkv = torch.zeros_like(qkv)
map_bytes = S.numel() * 4


def timeit(name, fn, nbytes, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:40s} {us:8.1f} us  {nbytes / us / 1e3:7.0f} GB/s (algorithmic bytes {nbytes / 1e6:.0f} MB)", flush=True)


rows = B * H * T
timeit("attn_scores (q k^T, masked)", lambda: K.attn_scores(q, k, valid, B, T, H, d, d ** -0.5, out=S), map_bytes + 2 * qkv.numel() * 2 // 3)
timeit("attn_map_loss kldiv", lambda: K.attn_map_loss(S, St, valid, valid, G, loss, B, T, H, 1, 1.0 / rows, 64.0), 2 * map_bytes + map_bytes // 2)
timeit("attn_map_loss mse", lambda: K.attn_map_loss(S, St, valid, valid, G, loss, B, T, H, 0, 1.0 / rows, 64.0), 2 * map_bytes + map_bytes // 2)
timeit("attn_scores_bwd dq = dS k", lambda: K.attn_scores_bwd(G, k, dqkv[:, :E], B, T, H, d, 0.125, trans=0), map_bytes // 2 + 3 * qkv.numel() * 2 // 3)
timeit("attn_scores_bwd dk = dS^T q", lambda: K.attn_scores_bwd(G, q, dqkv[:, E:2 * E], B, T, H, d, 0.125, trans=1), map_bytes // 2 + 3 * qkv.numel() * 2 // 3)
