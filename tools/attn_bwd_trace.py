"""Phase trace of one CTA of the tcgen05 attention backward (clock64 stamps of compute thread 0 and the MMA thread of
CTA (0, 0, 0)): FHB_LIB=<trace build> python tools/attn_bwd_trace.py [d T drop]
The trace build: tools/build_trace_lib.sh (attention_bwd_tc.cu with -DFHB_BWD_TRACE -> fithubert_b200/build/libfhb_trace.so)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from fithubert_b200 import kernels as K, lib as L

d, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (40, 389)
drop = (123, 0.1) if (len(sys.argv) <= 3 or sys.argv[3] != "0") else None
B, H = 32, 12
dev, f16 = "cuda", torch.float16
qkv = torch.randn(B, T, 3 * H * d, device=dev).to(f16)
vt = torch.full((B,), T, device=dev, dtype=torch.int32)
out, lse = torch.empty(B * T, H * d, device=dev, dtype=f16), torch.empty(B, H, T, device=dev)
K.attn_fwd(qkv, vt, out, lse, B, T, H, d, d ** -0.5, drop=drop)
do = torch.randn(B, T, H * d, device=dev).to(f16)
dqkv, delta, ws = torch.empty_like(qkv), torch.empty(B, H, T, device=dev), torch.empty(B * T, H * d, device=dev)
for _ in range(3):
    K.attn_bwd(qkv, vt, out, do, lse, dqkv, delta, B, T, H, d, d ** -0.5, dq_ws=ws, drop=drop)
torch.cuda.synchronize()
host = (C.c_longlong * 2048)()
L.check(L.lib().fhb_attn_bwd_trace_read(host, 2048), "trace_read")
v = list(host)
t0 = v[2047]
nq = (T + 127) // 128
names_c = ["top", "bar", "s_full", "ld0", "mma_done", "math0", "ld1", "s_free", "math1", "p_full", "drain", "dr.bar1", "dr.ld+st", "dr.bar2"]
order = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 13, 10]
print(f"attention backward d={d} T={T} drop={drop}: lane 0 of each compute warp (quarter = warp & 3, quad = warp >> 2), clk since CTA start")
for w in range(16):
    print(f"warp {w}:")
    prev = t0
    for i in range(min(nq, 4)):
        row = []
        for j in order:
            x = v[w * 64 + 16 * i + j]
            if not x:
                continue
            row.append(f"{names_c[j]}={x - t0}(+{x - prev})")
            prev = x
        print(f"  tile {i}: " + "  ".join(row))
names_m = ["top", "qd_full", "s_free", "S issued", "p_full", "dV dK issued", "dq_free", "dQ issued", "mma_done"]
print("MMA thread:")
for i in range(min(nq, 4)):
    row = [f"{n}={v[1024 + 16 * i + j] - t0}" for j, n in enumerate(names_m) if v[1024 + 16 * i + j]]
    print(f" tile {i}: " + "  ".join(row))
