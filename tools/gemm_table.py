"""Per-shape table of every fhb_gemm launch in one cfg-2 distillation step (CUDA-event pair per launch, PDL off):
shape, operand majors, epilogue flags, launches, total / average time, TFLOP/s.  Run on the GPU box."""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import fithubert_b200 as F  # noqa: E402
from fithubert_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
cfg = bench.yaml_cfg()
step = F.W2V2Distil(cfg, device=dev)
step.configure_optimizers(total_steps=1000)
x, pm, lengths = bench.synth_batch(32, 249600, 1234)
xd = x.to(dev)


def one():
    step.optimizer.zero_grad()
    step.fused_forward_backward(xd, None, lengths, grad_scale=1.0)
    step.optimizer_step()


for _ in range(3):
    one()
steps = 5
os.environ["FHB_STREAMS"] = "0"  # one stream: each event pair brackets exactly one kernel
K.enable_gemm_timing(True)
for _ in range(steps):
    one()
torch.cuda.synchronize()
tab = defaultdict(lambda: [0.0, 0.0, 0])
order = {}
for i, (e0, e1, fl, shape, meta) in enumerate(K._GEMM_TIMING["events"]):
    key = (shape, meta)
    t = tab[key]
    t[0] += e0.elapsed_time(e1) * 1e3
    t[1] += fl
    t[2] += 1
    order.setdefault(key, i)
K.enable_gemm_timing(False)
tot = sum(t[0] for t in tab.values()) / steps
print(f"GEMM time per step {tot / 1e3:.3f} ms; rows sorted by total time; first = index of the first launch in the step")
print(f"{'rows':>9} {'n':>5} {'k':>6} {'aMN':>3} {'bMN':>3} {'flags':>6} {'splK':>4} {'ob':>4} {'cnt':>4} {'us/step':>9} {'avg us':>8} {'TF/s':>7} {'first':>5}")
for (shape, meta), t in sorted(tab.items(), key=lambda kv: -kv[1][0]):
    cnt = t[2] // steps
    print(f"{shape[0]:9d} {shape[1]:5d} {shape[2]:6d} {meta[0]:3d} {meta[1]:3d} {meta[2]:#6x} {meta[3]:4d} {meta[4]:4d} {cnt:4d} "
          f"{t[0] / steps:9.1f} {t[0] / t[2]:8.1f} {t[1] / t[0] / 1e6:7.1f} {order[(shape, meta)]:5d}")
