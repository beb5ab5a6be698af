"""Where an epilogue slab's time goes: clock64 stamps of CTA 0 (threads 0 and 255) through the GEMM epilogue.
Needs the library built with tracing:  NVCC_EXTRA=-DFHB_GEMM_TRACE python -m fithubert_b200.build --force
usage: python tools/gemm_trace.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fithubert_b200 import kernels as K, lib as L

dev = "cuda"
f16, f32 = torch.float16, torch.float32
PH = ["tmem ld (+ring wait)", "math + st.shared", "fence (+store drain, thread 0)", "block barrier", "TMA store issue (thread 0)", "to next slab top"]


def trace(name, M, N, Kd, **kw):
    x, w, b = (torch.randn(M, Kd, device=dev) * 0.5).half(), (torch.randn(N, Kd, device=dev) * 0.05).half(), torch.randn(N, device=dev)
    out_dt = kw.pop("out_dt", f16)
    res = kw.pop("res", None)
    r = None if res is None else (torch.randn(M, N, device=dev)).to(res)
    u = torch.empty(M, N, device=dev, dtype=f16) if kw.pop("dg", False) else None
    out = torch.empty(M, N, device=dev, dtype=out_dt)
    trace_fn(name, lambda: K.linear(x, w, b, residual=r, out=out, out_dtype=out_dt, dgelu_out=u, **kw))


def trace_fn(name, run):
    for _ in range(3):
        run()
    buf = torch.zeros(2 * 96 * 8 + 32 + 1024, device=dev, dtype=torch.int64)
    L.check(L.lib().fhb_gemm_set_trace_buffer(C.c_void_p(buf.data_ptr())), "set_trace")
    e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
    run(); run()
    e0.record(); run(); e1.record(); run(); e2.record()  # the buffer keeps the stamps of the LAST launch
    torch.cuda.synchronize()
    L.check(L.lib().fhb_gemm_set_trace_buffer(None), "set_trace")
    ev_us = (e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3)
    tl = buf[2 * 96 * 8:2 * 96 * 8 + 32].cpu().tolist()
    cta = buf[2 * 96 * 8 + 32:].view(512, 2).cpu()
    v = buf[:2 * 96 * 8].view(2, 96, 8).cpu()
    print(f"== {name}")
    live = cta[:, 0] > 0
    if int(live.sum()) > 0:
        st, en = cta[live, 0], cta[live, 1]
        t0 = int(st.min())
        print(f"  CTAs {int(live.sum())}: start min 0 / median {int(st.median()) - t0} / max {int(st.max()) - t0} ns; end min {int(en.min()) - t0} / median "
              f"{int(en.median()) - t0} / max {int(en.max()) - t0} ns; lifetime median {int((en - st).median())} ns; CUDA-event time of the two "
              f"back-to-back launches around it {ev_us[0]:.1f} / {ev_us[1]:.1f} us")
    if tl[0]:
        names = {0: "entry", 1: "setup done", 2: "pdl wait passed", 3: "all warps done"}
        for i in range(4):
            names.update({4 + 4 * i: f"tile{i} first stage landed", 5 + 4 * i: f"tile{i} MMAs issued", 6 + 4 * i: f"tile{i} accumulator ready",
                          7 + 4 * i: f"tile{i} epilogue done"})
        ev = sorted((x - tl[0], names[i]) for i, x in enumerate(tl) if x and i in names)
        print("  timeline of CTA 0 (clk since entry): " + "; ".join(f"{n} {c}" for c, n in ev))
    for who, label in ((0, "thread 0"), (1, "thread 255")):
        t = v[who]
        n = int((t[:, 0] > 0).sum())
        if n < 12:
            print(f"  {label}: only {n} slabs traced")
            continue
        lo, hi = 6, min(n - 2, 60)
        d = []
        for s_ in range(lo, hi):
            row = t[s_]
            nxt = t[s_ + 1][0]
            d.append([int(row[1] - row[0]), int(row[2] - row[1]), int(row[3] - row[2]), int(row[4] - row[3]), int(row[5] - row[4]),
                      int(nxt - row[5])])
        d = torch.tensor(d, dtype=torch.float64)
        med = d.median(0).values.tolist()
        tot = float((t[hi][0] - t[lo][0]) / (hi - lo))
        print(f"  {label}: {tot:7.0f} clk per slab over slabs {lo}..{hi}; median per phase: " +
              "; ".join(f"{p} {m:.0f}" for p, m in zip(PH, med)))
        # tile boundaries: slot 6 = tile top, slot 7 = accumulator ready
        tb = [(int(t[s_][7] - t[s_][6])) for s_ in range(lo, hi) if t[s_][6] > 0 and t[s_][7] > 0]
        if tb:
            print(f"      tile top -> accumulator ready (bias barrier + acc_full wait): median {sorted(tb)[len(tb) // 2]} clk over {len(tb)} tiles")


Mc = 32 * 49919
trace("conv1 plain 1597408x256x128 (bias only)", Mc, 256, 128)
trace("conv1 gelu + dgelu out", Mc, 256, 128, gelu=True, dg=True)
trace("conv2 gelu + dgelu out 798688x256x768", 32 * 24959, 256, 768, gelu=True, dg=True)
trace("student fc2 res32 out32 12448x480x480", 32 * 389, 480, 480, res=f32, out_dt=f32)
trace("teacher out_proj + res16 24928x768x768", 32 * 779, 768, 768, res=f16)
Ms = 32 * 389
trace("student 12448x480x480 bias only (fp16 out)", Ms, 480, 480)
trace("student qkv 12448x1440x480 bias", Ms, 1440, 480)
dy = (torch.randn(Ms, 480, device=dev) * 0.5).half()
w = (torch.randn(480, 480, device=dev) * 0.05).half()
u = torch.rand(Ms, 480, device=dev).half()
r32 = torch.randn(Ms, 480, device=dev)
o16, o32 = torch.empty(Ms, 480, device=dev, dtype=f16), torch.empty(Ms, 480, device=dev, dtype=f32)
trace_fn("student dgrad 12448x480x480 plain (B MN-major)", lambda: K.linear_dgrad(dy, w, out=o16))
trace_fn("student dgrad 12448x480x480 x aux", lambda: K.linear_dgrad(dy, w, mul_aux=u, out=o16))
trace_fn("student dgrad 12448x480x480 + res32 -> fp32", lambda: K.linear_dgrad(dy, w, residual=r32, out=o32, out_dtype=f32))
dyq = (torch.randn(Ms, 1440, device=dev) * 0.5).half()
wq = (torch.randn(1440, 480, device=dev) * 0.05).half()
trace_fn("student dgrad qkv 12448x480x1440 + res32 -> fp32", lambda: K.linear_dgrad(dyq, wq, residual=r32, out=o32, out_dtype=f32))
