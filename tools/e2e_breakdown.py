"""Where does the end-to-end step (host buffers, one synchronisation per step) spend its time?  Per step: wall time, CPU
time to enqueue it, time blocked in the loss read-back, and the GPU span between an event recorded at step entry and one
recorded after the optimizer.  Run on the GPU box."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import fithubert_b200 as F  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
step = F.W2V2Distil(bench.yaml_cfg(), device=dev)
step.configure_optimizers(total_steps=1000)
x, pm, lengths = bench.synth_batch(32, 249600, 1234, pin=True)
xd = x.to(dev)


def run(kind, n=10):
    rows = []
    for i in range(n + 3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        t0 = time.perf_counter()
        e0.record()
        if kind == "host":
            loss = step.training_step({"x": x, "padding_mask": pm})
        elif kind == "host_lengths":
            loss = step.training_step({"x": x, "padding_mask": None, "lengths": lengths})
        else:
            loss = step.training_step({"x": xd, "padding_mask": None, "lengths": lengths})
        e1.record()
        t1 = time.perf_counter()
        v = float(loss.detach())
        t2 = time.perf_counter()
        if i >= 3:
            rows.append((1e3 * (t2 - t0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), e0.elapsed_time(e1)))
    m = [sum(r[j] for r in rows) / len(rows) for j in range(4)]
    print(f"{kind:13s} wall {m[0]:7.3f} ms  enqueue {m[1]:7.3f}  blocked in read-back {m[2]:7.3f}  GPU span {m[3]:7.3f}  loss {v:.5f}")


for kind in ("device", "host_lengths", "host", "device", "host"):
    run(kind)
# bare H2D bandwidth of the 32 MB batch
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    xd.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D {x.numel() * 4 / 1e6:.1f} MB in {dt * 1e3:.3f} ms = {x.numel() * 4 / dt / 1e9:.1f} GB/s")
