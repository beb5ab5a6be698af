"""Run ONE kernel case a few times (for `ncu --set full`): python tools/one_kernel.py <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K

dev, bf = "cuda", torch.bfloat16
case = sys.argv[1]
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).to(bf)
if case == "attn64":
    B, T, H, d = 32, 779, 12, 64
    qkv = rnd(B, T, 3 * H * d)
    vt = torch.tensor([T - 3 * i for i in range(B)], device=dev, dtype=torch.int32)
    out, lse = torch.empty(B * T, H * d, device=dev, dtype=bf), torch.empty(B, H, T, device=dev)
    fn = lambda: K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5)
elif case in ("attn40bwd", "attn40bwd_drop"):
    B, T, H, d = 32, 389, 12, 40
    qkv = rnd(B, T, 3 * H * d)
    vt = torch.tensor([T - 3 * i for i in range(B)], device=dev, dtype=torch.int32)
    out, lse = torch.empty(B * T, H * d, device=dev, dtype=bf), torch.empty(B, H, T, device=dev)
    K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5)
    do, dqkv, delta = rnd(B, T, H * d), torch.empty_like(qkv), torch.empty(B, H, T, device=dev)
    ws = torch.empty(B * T, H * d, device=dev)
    drop = (123, 0.1) if case.endswith("drop") else None
    fn = lambda: K.attn_bwd(qkv, vt, out, do, lse, dqkv, delta, B, T, H, d, d ** -0.5, dq_ws=ws, drop=drop)
elif case == "conv1_gelu":
    M, N, Kd = 32 * 49919, 256, 128
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out, gelu=True)
elif case == "conv1_plain":
    M, N, Kd = 32 * 49919, 256, 128
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out)
elif case == "sfc":
    M, N, Kd = 32 * 389, 480, 480
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out)
elif case == "tqkv":
    M, N, Kd = 32 * 779, 2304, 768
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out)
for _ in range(3):
    fn()
torch.cuda.synchronize()
