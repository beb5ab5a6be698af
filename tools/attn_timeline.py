"""Phase timeline of one CTA of the tcgen05 attention forward (clock64 stamps): python tools/attn_timeline.py [d T]
Needs the library built with -DFHB_ATTN_TIMELINE (NVCC_EXTRA=-DFHB_ATTN_TIMELINE python -m fithubert_b200.build --force)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from fithubert_b200 import kernels as K, lib as L

d, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 779)
B, H = 32, 12
dev, bf = "cuda", torch.bfloat16
qkv = torch.randn(B, T, 3 * H * d, device=dev).to(bf)
out, lse = torch.empty(B * T, H * d, device=dev, dtype=bf), torch.empty(B, H, T, device=dev)
for _ in range(3):
    K.attn_fwd(qkv, None, out, lse, B, T, H, d, d ** -0.5)
buf = torch.zeros(1024, device=dev, dtype=torch.int64)
L.lib().fhb_attn_fwd_set_profile_buffer.restype = C.c_int
L.check(L.lib().fhb_attn_fwd_set_profile_buffer(C.c_void_p(buf.data_ptr())), "set_profile")
K.attn_fwd(qkv, None, out, lse, B, T, H, d, d ** -0.5)
torch.cuda.synchronize()
L.check(L.lib().fhb_attn_fwd_set_profile_buffer(None), "set_profile")
v = buf.cpu().tolist()
nt = (T + 127) // 128
t0 = min(x for x in v if x > 0)
names_s = ["wait s_full", "s_full seen", "S in regs", "max done", "bar1 passed", "pre o_done", "o_done seen", "bar2 passed", "exp+STS done", "p_full arrived"]
print(f"softmax warp 0 (clk since CTA start {v[15] - t0}):")
for j in range(nt):
    row = [v[16 * j + i] - t0 if v[16 * j + i] else None for i in range(10)]
    print(f" tile {j}: " + "  ".join(f"{n}={r}" for n, r in zip(names_s, row)))
    if j:
        prev = v[16 * (j - 1) + 9]
        print(f"    deltas: " + " ".join(str((v[16 * j + i] - (v[16 * j + i - 1] if i else prev)) if v[16 * j + i] and (v[16 * j + i - 1] if i else prev) else None) for i in range(10)))
print("final o_done seen", v[16 * nt] - t0)
names_m = ["loop top", "kv_full seen", "s_free seen", "S issued", "p_full seen", "PV issued+commits"]
print("TMA/MMA warp:")
for j in range(nt):
    row = [v[512 + 16 * j + i] - t0 if v[512 + 16 * j + i] else None for i in range(6)]
    print(f" tile {j}: " + "  ".join(f"{n}={r}" for n, r in zip(names_m, row)))
