"""torchrun tools/nccl_check.py : all-reduce bandwidth of the flat gradient buffer size + transport info."""
import os, time, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
x = torch.ones(31_168_352, device="cuda")
for n_b in (1, 6):
    step = -(-x.numel() // n_b)
    for _ in range(3):
        for a in range(0, x.numel(), step):
            dist.all_reduce(x[a:a + step])
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        for a in range(0, x.numel(), step):
            dist.all_reduce(x[a:a + step])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"buckets={n_b}: {ms:.3f} ms per 124.7 MB all-reduce, algbw {x.numel()*4/ms/1e6:.1f} GB/s", flush=True)
# side-stream pattern used by GradAllReduce
s = torch.cuda.Stream()
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(10):
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for a in range(0, x.numel(), step):
            dist.all_reduce(x[a:a + step])
    torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
if rank == 0:
    print(f"side-stream pattern: {(time.perf_counter()-t0)*100:.3f} ms per step", flush=True)
dist.destroy_process_group()
