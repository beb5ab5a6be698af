"""Isolated timings of the epilogue-heavy GEMM shapes of the step (CUDA events, 30 reps after 5 warm-ups; inputs cycle
through 4 buffer sets so that L2 residency resembles the step's).  usage: python tools/epi_probe.py [tag]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K

dev = "cuda"
f16, f32 = torch.float16, torch.float32


def rnd(*s, sc=1.0, dt=f16):
    return (torch.randn(*s, device=dev) * sc).to(dt)


def timeit(name, fns, flops, reps=30):
    for i in range(5):
        fns[i % len(fns)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for i in range(reps):
        fns[i % len(fns)]()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print(f"{name:44s} {us:8.1f} us {flops / us / 1e6:8.1f} TF/s", flush=True)


def case(name, M, N, Kd, res=None, out_dt=f16, drop=None, gelu=False, dg=False, sets=4):
    fns = []
    for _ in range(sets):
        x, w, b = rnd(M, Kd), rnd(N, Kd, sc=0.05), torch.randn(N, device=dev)
        r = None if res is None else rnd(M, N, dt=res)
        out = torch.empty(M, N, device=dev, dtype=out_dt)
        u = torch.empty(M, N, device=dev, dtype=f16) if dg else None
        fns.append(lambda x=x, w=w, b=b, r=r, out=out, u=u: K.linear(x, w, b, residual=r, out=out, out_dtype=out_dt, drop=drop,
                                                                     gelu=gelu, dgelu_out=u))
    timeit(name, fns, 2.0 * M * N * Kd)


Mt, Ms = 32 * 779, 32 * 389
case("S plain            12448x480x480", Ms, 480, 480)
case("S qkv              12448x1440x480", Ms, 1440, 480)
case("S fc1 gelu+dgelu   12448x480x480", Ms, 480, 480, gelu=True, dg=True)
case("S res16 out16      12448x480x480", Ms, 480, 480, res=f16)
case("S res32 out32      12448x480x480", Ms, 480, 480, res=f32, out_dt=f32)
case("S res32 out32 drop 12448x480x480", Ms, 480, 480, res=f32, out_dt=f32, drop=(1234, 0.1))
case("S out32 (no res)   12448x480x480", Ms, 480, 480, out_dt=f32)
case("T qkv              24928x2304x768", Mt, 2304, 768)
case("T out+res16        24928x768x768", Mt, 768, 768, res=f16)
case("T out (no res)     24928x768x768", Mt, 768, 768)
case("T fc2+res16        24928x768x3072", Mt, 768, 3072, res=f16)
if len(sys.argv) > 1 and sys.argv[1] == "conv":
    Mc = 32 * 49919
    print("-- student conv1 shape (1.6 M rows, K = 128): where the 0x212 forward epilogue's time goes")
    case("C1 plain (no bias)  1597408x256x128", Mc, 256, 128, sets=2)
    case("C1 gelu             1597408x256x128", Mc, 256, 128, gelu=True, sets=2)
    case("C1 gelu + dgelu out 1597408x256x128", Mc, 256, 128, gelu=True, dg=True, sets=2)
    case("C2 gelu + dgelu out 798688x256x768", 32 * 24959, 256, 768, gelu=True, dg=True, sets=2)
    case("C2 plain            798688x256x768", 32 * 24959, 256, 768, sets=2)
