#!/bin/bash
# bench at N GPUs exactly as the driver launches it (no N=1 leg, no reference arm).  usage: tools/gpu_scale_only.sh N TAG
N=${1:-8}; TAG=${2:-scale}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
tail -n 3 gpurun_out/${TAG}_n$N.err | cut -c1-300
python tools/print_bench.py gpurun_out/${TAG}_n$N.json
