"""Where the HOST time of one step goes: cProfile over a few device steps of bench.py's cfg-2 workload (no synchronisation
inside the profiled region).  usage: python tools/host_profile.py [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import torch
import bench

steps = 5
ns = {}
# reuse bench.main's setup through its --profile path: monkeypatch the early return to hand back dev_step
src = open(bench.__file__).read()
src = src.replace("    if args.profile:\n        for _ in range(2):\n            dev_step()\n        torch.cuda.synchronize()\n",
                  "    if args.profile:\n        for _ in range(3):\n            dev_step()\n        torch.cuda.synchronize()\n        globals()['_DEV_STEP'] = dev_step\n")
g = {"__name__": "bench_host_profile", "__file__": bench.__file__}
exec(compile(src, bench.__file__, "exec"), g)
sys.argv = [bench.__file__, "--profile"]
g["main"]()
dev_step = g["_DEV_STEP"]
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(steps):
    dev_step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / steps:.2f} ms per step; with the device drained {1e3 * (t2 - t0) / steps:.2f} ms per step")
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(steps):
    dev_step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
