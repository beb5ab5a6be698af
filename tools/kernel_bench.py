"""Per-kernel micro-benchmarks at cfg-2 shapes (CUDA events, 20 reps after 3 warm-ups).
usage: python tools/kernel_bench.py [filter-substring ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K, lib as L

dev = "cuda"
bf = torch.float16  # the library's 16-bit storage format
sel = sys.argv[1:]


def rnd(*s, sc=1.0):
    return (torch.randn(*s, device=dev) * sc).to(bf)


def timeit(name, fn, flops=0.0, bytes_=0.0, reps=20):
    if sel and not any(x in name for x in sel):
        return
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    extra = ""
    if flops:
        extra += f" {flops / us / 1e6:8.1f} TF/s"
    if bytes_:
        extra += f" {bytes_ / us / 1e3:8.1f} GB/s"
    print(f"{name:46s} {us:9.1f} us{extra}", flush=True)


def gemm_case(name, M, N, Kd, **kw):
    x, w = rnd(M, Kd), rnd(N, Kd, sc=0.05)
    bias = torch.randn(N, device=dev)
    args = {}
    if kw.get("res"):
        args["residual"] = rnd(M, N)
    if kw.get("gelu"):
        args["gelu"] = True
    if kw.get("pre"):
        args["preact_out"] = torch.empty(M, N, device=dev, dtype=bf)
    if kw.get("dg"):
        args["dgelu_out"] = torch.empty(M, N, device=dev, dtype=bf)
    out = torch.empty(M, N, device=dev, dtype=bf)
    timeit(name, lambda: K.linear(x, w, bias, out=out, **args), flops=2.0 * M * N * Kd)


def dgrad_case(name, M, N, Kd, **kw):
    dy, w = rnd(M, N), rnd(N, Kd, sc=0.05)
    args = {}
    if kw.get("res"):
        args["residual"] = rnd(M, Kd)
    if kw.get("aux"):
        args["mul_aux"] = rnd(M, Kd)
    out = torch.empty(M, Kd, device=dev, dtype=bf)
    timeit(name, lambda: K.linear_dgrad(dy, w, out=out, **args), flops=2.0 * M * N * Kd)


def wgrad_case(name, M, N, Kd):
    dy, x = rnd(M, N, sc=0.1), rnd(M, Kd)
    out = torch.zeros(N, Kd, device=dev)
    timeit(name, lambda: K.linear_wgrad(dy, x, out=out, accumulate=True), flops=2.0 * M * N * Kd)


Mt, Ms = 32 * 779, 32 * 389
gemm_case("T qkv      24928x2304x768", Mt, 2304, 768)
gemm_case("T out+res  24928x768x768", Mt, 768, 768, res=True)
gemm_case("T fc1+gelu 24928x3072x768", Mt, 3072, 768, gelu=True)
gemm_case("T fc2+res  24928x768x3072", Mt, 768, 3072, res=True)
gemm_case("S qkv      12448x1440x480", Ms, 1440, 480)
gemm_case("S out+res  12448x480x480", Ms, 480, 480, res=True)
gemm_case("S fc1+gelu+dg 12448x480x480", Ms, 480, 480, gelu=True, dg=True)
gemm_case("S fc1 plain 12448x480x480", Ms, 480, 480)
gemm_case("S conv1 1.6Mx256x128 plain", 32 * 49919, 256, 128)
gemm_case("S conv1 1.6Mx256x128 gelu", 32 * 49919, 256, 128, gelu=True)
gemm_case("S conv1 1.6Mx256x128 gelu+pre", 32 * 49919, 256, 128, gelu=True, pre=True)
gemm_case("S conv1 1.6Mx256x128 gelu+dg", 32 * 49919, 256, 128, gelu=True, dg=True)
gemm_case("S conv6 100Kx512x256 gelu+dg", 32 * 3119, 512, 256, gelu=True, dg=True)
dgrad_case("S dgrad fc2*aux 12448x480->480", Ms, 480, 480, aux=True)
dgrad_case("S dgrad fc1+res 12448x480->480", Ms, 480, 480, res=True)
dgrad_case("S dgrad qkv+res 12448x1440->480", Ms, 1440, 480, res=True)
dgrad_case("S dgrad conv1 1.6Mx256->128", 32 * 49919, 256, 128)
dgrad_case("S dgrad conv2e*aux 400Kx512->256", 32 * 12480, 512, 256, aux=True)
wgrad_case("S wgrad fc 12448: 480x480", Ms, 480, 480)
wgrad_case("S wgrad qkv 12448: 1440x480", Ms, 1440, 480)
wgrad_case("S wgrad conv1 1.6M: 256x128", 32 * 49919, 256, 128)

for (d, T, H) in ((64, 779, 12), (40, 389, 12)):
    B = 32
    qkv = rnd(B, T, 3 * H * d)
    vt = torch.tensor([T - 3 * i for i in range(B)], device=dev, dtype=torch.int32)
    out = torch.empty(B * T, H * d, device=dev, dtype=bf)
    lse = torch.empty(B, H, T, device=dev)
    timeit(f"attn fwd d={d} T={T}", lambda: K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5), flops=4.0 * B * H * T * T * d)
    do = rnd(B, T, H * d)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, T, device=dev)
    ws = torch.empty(B * T, H * d, device=dev)
    timeit(f"attn bwd d={d} T={T} (tcgen05)", lambda: K.attn_bwd(qkv, vt, out, do, lse, dqkv, delta, B, T, H, d, d ** -0.5, dq_ws=ws),
           flops=10.0 * B * H * T * T * d)
    timeit(f"attn bwd d={d} T={T} +dropout", lambda: K.attn_bwd(qkv, vt, out, do, lse, dqkv, delta, B, T, H, d, d ** -0.5, dq_ws=ws,
                                                              drop=(123, 0.1)), flops=10.0 * B * H * T * T * d)

for (rows, C) in ((Mt, 768), (Ms, 480)):
    x = rnd(rows, C)
    g, b = torch.randn(C, device=dev), torch.randn(C, device=dev)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    timeit(f"LN fwd {rows}x{C}", lambda: K.layernorm_fwd(x, g, b, y, mean, rstd), bytes_=rows * C * 4.0)
    dg, db, ds = torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    timeit(f"LN bwd {rows}x{C} (+dxsum)", lambda: K.layernorm_bwd(y, x, g, mean, rstd, y, dg, db, dxsum=ds), bytes_=rows * C * 6.0)
    timeit(f"colsum {rows}x{C}", lambda: K.colsum(x, ds), bytes_=rows * C * 2.0)

for (C, nm) in ((512, "teacher"), (128, "student")):
    B, Ld = 32, 249600
    T0 = (Ld - 10) // 5 + 1
    wave = 0.1 * torch.randn(B, Ld, device=dev)
    w = torch.randn(C, 1, 10, device=dev) * 0.45
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    stat = torch.empty(B, 65, device=dev, dtype=torch.float64)
    mean, rstd = torch.empty(B, C, device=dev), torch.empty(B, C, device=dev)
    out = torch.empty(B, T0, C, device=dev, dtype=bf)
    timeit(f"conv0 fwd {nm} C={C}", lambda: K.conv0_fwd(wave, w, g, b, T0, stat, mean, rstd, out), bytes_=B * T0 * C * 2.0 + B * Ld * 4.0)
    if C == 128:
        acc = torch.empty(B, C, 12, device=dev)
        dw, dgm, dbt = torch.zeros(C, 10, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
        timeit(f"conv0 bwd {nm} C={C}", lambda: K.conv0_bwd(wave, w, g, b, T0, stat, mean, rstd, out, acc, dw, dgm, dbt),
               bytes_=B * T0 * C * 2.0 + B * Ld * 4.0)
