#!/bin/bash
# N-GPU A/B of the overlapped gradient exchange: SMs left to NCCL (FHB_COMM_SMS = NCCL_MAX_CTAS) vs no overlap.  usage: N TAG
N=${1:-2}; TAG=${2:-commsms}
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
port=29540
for rep in 1 2; do
  for vv in "0 8" "1 2" "1 4" "1 8"; do
    set -- $vv
    v=$1; sms=$2
    port=$((port + 1))
    echo "=== FHB_EARLY_REDUCE=$v FHB_COMM_SMS=$sms rep $rep" >> $LOG
    FHB_EARLY_REDUCE=$v FHB_COMM_SMS=$sms timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_v${v}_s${sms}_$rep.json 2> gpurun_out/${TAG}_err.log
    python - >> $LOG 2>&1 <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_v${v}_s${sms}_$rep.json").read().strip().splitlines()[-1])
print("ms_per_step %.3f value %.0f e2e %.0f comm_exposed_ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("comm_exposed_ms", -1)))
PY
  done
done
cat $LOG
