"""torchrun tools/ddp_probe.py : where does a data-parallel step spend its time (device events + host clock)."""
import os, sys, time, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
import fithubert_b200 as F
torch.manual_seed(0)
cfg = bench.yaml_cfg()
step = F.W2V2Distil(cfg, device=dev)
step.configure_optimizers(total_steps=1000)
x_host, pm_host, lengths = bench.synth_batch(32, 249600, 1234 + rank, pin=True)
x = x_host.to(dev)
def phase(name, fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"[rank {rank}] {name:28s} device {e0.elapsed_time(e1)/n:8.2f} ms  host-enqueue {(t1-t0)/n*1e3:8.2f} ms  wall {(t2-t0)/n*1e3:8.2f} ms", flush=True)
def fb():
    step.optimizer.zero_grad()
    step.fused_forward_backward(x, None, lengths)
def fb_red():
    fb()
    _, _, G = step.student_model.engine_state(True)
    step.reducer.reduce_all(G.flat); step.reducer.wait()
def full():
    fb(); step.optimizer_step()
phase("fwd+bwd", fb)
phase("fwd+bwd+allreduce", fb_red)
phase("full step", full)
phase("full step again", full)
if world > 1:
    dist.destroy_process_group()
