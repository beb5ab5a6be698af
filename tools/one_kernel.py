"""Run ONE kernel case a few times (for `ncu --set full`): python tools/one_kernel.py <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K

dev, bf = "cuda", torch.float16
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
elif case == "sfc_res":
    M, N, Kd = 32 * 389, 480, 480
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    res = rnd(M, N)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out, residual=res)
elif case == "tqkv":
    M, N, Kd = 32 * 779, 2304, 768
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=bf)
    fn = lambda: K.linear(x, w, b, out=out)
elif case == "pc_fin_bwd":
    B, T, Cd, G, cp, dl = 32, 779, 480, 16, 32, 4
    R = (T + dl - 1) // dl
    Tp = dl * R + 132
    dy, h = rnd(B * T, Cd), rnd(B * T, Cd)
    conv = rnd(B * R, G * dl * cp)
    bias, gamma = torch.randn(Cd, device=dev), torch.randn(Cd, device=dev)
    mean, rstd = torch.randn(B * T, device=dev), torch.rand(B * T, device=dev) + 0.5
    dh = torch.empty(B * T, Cd, device=dev, dtype=bf)
    dcg = torch.zeros(B * G, Tp, cp, device=dev, dtype=bf)
    dg, db, dbi = (torch.zeros(Cd, device=dev) for _ in range(3))
    fn = lambda: K.posconv_finish_bwd(dy, h, conv, bias, gamma, mean, rstd, dh, dcg, dg, db, dbi, B, T, Cd, G, cp, 63, Tp, delta=dl)
elif case == "ln_bwd":
    rows, Cd = 32 * 389, 480
    x, dy, dy2 = rnd(rows, Cd), rnd(rows, Cd), rnd(rows, Cd)
    g = torch.randn(Cd, device=dev)
    mean, rstd = torch.randn(rows, device=dev), torch.rand(rows, device=dev) + 0.5
    dx, dxd = torch.empty_like(x), torch.empty_like(x)
    dg, db, ds = (torch.zeros(Cd, device=dev) for _ in range(3))
    fn = lambda: K.layernorm_bwd(dy, x, g, mean, rstd, dx, dg, db, dxsum=ds, dy2=dy2, dx_drop=dxd, drop=(77, 0.1))
elif case == "conv0_bwd":
    B, Ld, Cd = 32, 249600, 128
    T0 = (Ld - 10) // 5 + 1
    wave = 0.1 * torch.randn(B, Ld, device=dev)
    w = torch.randn(Cd, 1, 10, device=dev) * 0.45
    g, b = torch.ones(Cd, device=dev), torch.zeros(Cd, device=dev)
    stat = torch.empty(B, 65, device=dev, dtype=torch.float64)
    mean, rstd = torch.empty(B, Cd, device=dev), torch.empty(B, Cd, device=dev)
    out = torch.empty(B, T0, Cd, device=dev, dtype=bf)
    K.conv0_fwd(wave, w, g, b, T0, stat, mean, rstd, out)
    acc = torch.empty(B, Cd, 12, device=dev)
    dw, dgm, dbt = torch.zeros(Cd, 10, device=dev), torch.zeros(Cd, device=dev), torch.zeros(Cd, device=dev)
    fn = lambda: K.conv0_bwd(wave, w, g, b, T0, stat, mean, rstd, out, acc, dw, dgm, dbt, dy_is_dz=True)
elif case in ("conv0_fwd_t", "conv0_fwd_s"):
    B, Ld, Cd = 32, 249600, (512 if case.endswith("t") else 128)
    T0 = (Ld - 10) // 5 + 1
    wave = 0.1 * torch.randn(B, Ld, device=dev)
    w = torch.randn(Cd, 1, 10, device=dev) * 0.45
    g, b = torch.ones(Cd, device=dev), torch.zeros(Cd, device=dev)
    stat = torch.empty(B, 65, device=dev, dtype=torch.float64)
    mean, rstd = torch.empty(B, Cd, device=dev), torch.empty(B, Cd, device=dev)
    out = torch.empty(B, T0, Cd, device=dev, dtype=bf)
    gp = torch.empty(B, T0, Cd, device=dev, dtype=bf) if Cd == 128 else None
    fn = lambda: K.conv0_fwd(wave, w, g, b, T0, stat, mean, rstd, out, gp_out=gp)
elif case == "sfc_res32":
    M, N, Kd = 32 * 389, 480, 480
    x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32)
    fn = lambda: K.linear(x, w, b, out=out, residual=res, out_dtype=torch.float32)
elif case == "ln_fwd_t":
    rows, Cd = 32 * 779, 768
    x = rnd(rows, Cd); y = torch.empty_like(x)
    g, b = torch.randn(Cd, device=dev), torch.randn(Cd, device=dev)
    fn = lambda: K.layernorm_fwd(x, g, b, y)
elif case == "ln_fwd32_s":
    rows, Cd = 32 * 389, 480
    x = torch.randn(rows, Cd, device=dev); y = torch.empty(rows, Cd, device=dev, dtype=bf); y32 = torch.empty_like(x)
    g, b = torch.randn(Cd, device=dev), torch.randn(Cd, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    fn = lambda: K.layernorm_fwd32(x, g, b, y, y32, mean, rstd)
for _ in range(3):
    fn()
torch.cuda.synchronize()
