#!/bin/bash
# N-GPU bench with the two-phase gradient exchange on.  usage: tools/gpu_scale_early.sh N TAG
N=${1:-8}; TAG=${2:-scale_early}
mkdir -p gpurun_out
FHB_EARLY_REDUCE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
tail -n 3 gpurun_out/${TAG}_n$N.err | cut -c1-300
python tools/print_bench.py gpurun_out/${TAG}_n$N.json
