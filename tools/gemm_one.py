"""Launch a handful of representative GEMMs (for ncu --set full captures)."""
import sys
import torch
sys.path.insert(0, ".")
from fithubert_b200 import kernels as K
torch.manual_seed(0)
shapes = {"tqkv": (24928, 2304, 768, False), "tfc1": (24928, 3072, 768, True), "sqkv": (12448, 1440, 480, False)}
M, N, Kd, gelu = shapes[sys.argv[1] if len(sys.argv) > 1 else "tqkv"]
x = torch.randn(M, Kd, device="cuda").bfloat16()
w = (torch.randn(N, Kd, device="cuda") * 0.05).bfloat16()
b = torch.randn(N, device="cuda")
for _ in range(3):
    K.linear(x, w, b, gelu=gelu)
torch.cuda.synchronize()
