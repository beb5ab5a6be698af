#!/bin/bash
# bench at N GPUs exactly as the driver launches it.  usage: tools/gpu_scale.sh N TAG
N=${1:-2}; TAG=${2:-scale}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
tail -3 gpurun_out/${TAG}_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/${TAG}_ref_n$N.json 2> gpurun_out/${TAG}_ref_n$N.err
python - <<PY
import json
for f in ("gpurun_out/${TAG}_n1.json", "gpurun_out/${TAG}_n$N.json", "gpurun_out/${TAG}_ref_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), round(d["value"], 1), d["unit"], "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1))
    except Exception as e:
        print(f, "unreadable:", e)
PY
