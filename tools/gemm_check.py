"""GPU check of fhb_gemm variants against torch (fp32 matmul of the same bf16 inputs).
Usage: python tools/gemm_check.py <case>   (each case in its own process: a device trap kills only it)"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from fithubert_b200 import kernels as K, lib as L

dev = "cuda"
bf = torch.bfloat16


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9))


def report(name, got, ref, tol=1e-2):
    e = rel(got, ref)
    print(f"{'OK ' if e < tol else 'BAD'} {name}: rel={e:.3e}", flush=True)
    return e < tol


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(bf)


def case_kk():
    ok = True
    for (M, N, Kd) in [(128, 64, 64), (256, 256, 128), (1000, 480, 480), (777, 768, 768), (300, 1440, 480), (129, 48, 4096)]:
        x, w = rnd(M, Kd), rnd(N, Kd, scale=0.05)
        y = K.linear(x, w)
        ok &= report(f"kk {M}x{N}x{Kd}", y, x.float() @ w.float().t())
    return ok


def case_epi():
    ok = True
    M, N, Kd = 600, 480, 512
    x, w, b = rnd(M, Kd), rnd(N, Kd, scale=0.05), torch.randn(N, device=dev)
    res = rnd(M, N)
    ref = x.float() @ w.float().t() + b
    ok &= report("bias", K.linear(x, w, b), ref)
    ok &= report("bias+gelu", K.linear(x, w, b, gelu=True), F.gelu(ref))
    pre = torch.empty(M, N, device=dev, dtype=bf)
    y = K.linear(x, w, b, gelu=True, preact_out=pre)
    ok &= report("gelu+preact y", y, F.gelu(ref))
    ok &= report("gelu+preact pre", pre, ref)
    ok &= report("bias+res", K.linear(x, w, b, residual=res), ref + res.float())
    ok &= report("f32 out", K.linear(x, w, b, out_dtype=torch.float32), ref, tol=1e-4)
    rv = torch.tensor([100, 250, 0], device=dev, dtype=torch.int32)
    y = K.linear(x, w, b, row_valid=rv, rows_per_batch=200)
    r3 = ref.view(3, 200, N).clone()
    for i, v in enumerate([100, 200, 0]):
        r3[i, v:] = 0
    ok &= report("rowzero", y, r3.view(M, N))
    return ok


def case_conv():
    """k=3,s=2 and k=2,s=2 convs as overlapped / reshaped views."""
    ok = True
    B, T, Cin, Cout = 3, 1001, 256, 512
    x = rnd(B, T, Cin)
    w = rnd(Cout, Cin, 3, scale=0.05)
    To = (T - 3) // 2 + 1
    ref = F.gelu(F.conv1d(x.float().transpose(1, 2), w.float(), stride=2)).transpose(1, 2)
    w2 = w.permute(0, 2, 1).reshape(Cout, 3 * Cin).contiguous()
    y = torch.empty(B, To, Cout, device=dev, dtype=bf)
    a3 = L.tensor3(data_ptr=x.data_ptr(), dim=(3 * Cin, To, B), stride=(2 * Cin, T * Cin))
    K.gemm_raw(a3, L.tensor3(w2), y, To, Cout, 3 * Cin, num_ob=B, a_coord=(0, 1, 0, 0), d_ld=Cout,
               d_hi_stride=To * Cout, flags=L.EPI_GELU)
    ok &= report("conv k3s2 overlapped-stride TMA", y, ref)
    return ok


def case_dgrad():
    ok = True
    for (M, N, Kd) in [(500, 480, 480), (1000, 768, 3072), (300, 960, 480), (260, 64, 128)]:
        dy, w = rnd(M, N), rnd(N, Kd, scale=0.05)
        ok &= report(f"dgrad {M}x{N}x{Kd}", K.linear_dgrad(dy, w), dy.float() @ w.float())
    M, N, Kd = 500, 480, 480
    dy, w, u, r = rnd(M, N), rnd(N, Kd, scale=0.05), rnd(M, Kd), rnd(M, Kd)
    uu = u.float().requires_grad_(True)
    F.gelu(uu).backward(dy.float() @ w.float())
    ok &= report("dgrad*dgelu+res", K.linear_dgrad(dy, w, dgelu_of=u, residual=r), uu.grad + r.float())
    return ok


def case_wgrad():
    ok = True
    for (M, N, Kd) in [(256, 128, 64), (5000, 480, 480), (3000, 768, 480), (4096, 480, 960), (1111, 3072, 768)]:
        dy, x = rnd(M, N, scale=0.1), rnd(M, Kd)
        ok &= report(f"wgrad {M}x{N}x{Kd}", K.linear_wgrad(dy, x), dy.float().t() @ x.float(), tol=2e-3)
    return ok


def case_perf():
    for (M, N, Kd, tag) in [(24928, 3072, 768, "T fc1"), (24928, 768, 3072, "T fc2"), (24928, 2304, 768, "T qkv"),
                            (12448, 1440, 480, "S qkv"), (12448, 480, 480, "S fc"), (32 * 24959, 512, 1536, "T conv1")]:
        x, w = rnd(M, Kd), rnd(N, Kd, scale=0.05)
        b = torch.randn(N, device=dev)
        for _ in range(3):
            K.linear(x, w, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            K.linear(x, w, b)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        t0 = torch.cuda.Event(True); t1 = torch.cuda.Event(True)
        for _ in range(3):
            F.linear(x, w)
        t0.record()
        for _ in range(10):
            F.linear(x, w)
        t1.record()
        torch.cuda.synchronize()
        ms_t = t0.elapsed_time(t1) / 10
        fl = 2.0 * M * N * Kd
        print(f"perf {tag} {M}x{N}x{Kd}: fhb {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s | cublas {ms_t:.3f} ms {fl / ms_t / 1e9:.0f} TF/s", flush=True)
    return True


if __name__ == "__main__":
    torch.manual_seed(0)
    name = sys.argv[1]
    t = time.time()
    ok = globals()["case_" + name]()
    torch.cuda.synchronize()
    print(f"case {name}: {'PASS' if ok else 'FAIL'} ({time.time() - t:.1f}s)", flush=True)
    sys.exit(0 if ok else 1)
