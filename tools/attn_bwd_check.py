"""Focused check of the tcgen05 attention backward (timeout-wrapped on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K
def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
torch.manual_seed(3)
for (d, T, B, H, short) in [(64, 128, 1, 1, 0), (40, 128, 1, 1, 0), (40, 389, 2, 3, 37), (64, 779, 2, 3, 300), (40, 130, 2, 2, 129)]:
    qkv = torch.randn(B, T, 3 * H * d, device="cuda").bfloat16()
    vt = torch.tensor([T] + [max(1, T - short)] * (B - 1), device="cuda", dtype=torch.int32)
    out, lse = torch.empty(B * T, H * d, device="cuda", dtype=torch.bfloat16), torch.empty(B, H, T, device="cuda")
    K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5)
    q3 = qkv.float().requires_grad_(True)
    q, k, v = (t.reshape(B, T, H, d).transpose(1, 2) for t in q3.chunk(3, dim=-1))
    mask = (torch.arange(T, device="cuda")[None] >= vt[:, None])[:, None, None, :]
    s = ((q @ k.transpose(-1, -2)) * d ** -0.5).masked_fill(mask, float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, T, H * d)
    do = torch.randn(B, T, H * d, device="cuda").bfloat16()
    ref.backward(do.float())
    delta = torch.empty(B, H, T, device="cuda")
    dq2 = torch.full_like(qkv, float("nan"))
    K.attn_bwd(qkv, vt, out, do, lse, dq2, delta, B, T, H, d, d ** -0.5, dq_ws=torch.empty(B * T, H * d, device="cuda"))
    torch.cuda.synchronize()
    g = q3.grad
    E = H * d
    print(f"d={d} T={T} B={B} H={H}: dq {rel(dq2[..., :E], g[..., :E]):.4f} dk {rel(dq2[..., E:2*E], g[..., E:2*E]):.4f} "
          f"dv {rel(dq2[..., 2*E:], g[..., 2*E:]):.4f} nan={int(torch.isnan(dq2.float()).sum())}", flush=True)
