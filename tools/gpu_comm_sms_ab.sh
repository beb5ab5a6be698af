#!/bin/bash
# N-GPU A/B of the overlapped gradient exchange.  Each case = "EARLY_REDUCE COMM_SMS WINDOW": exchange under the backward or
# after it, SMs left to NCCL (= NCCL_MAX_CTAS), reservation only for one layer after each hand-over or for the whole backward.
# usage: tools/gpu_comm_sms_ab.sh N TAG
N=${1:-2}; TAG=${2:-commsms}
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
port=29540
for rep in 1 2; do
  for vv in "0 8 0" "1 8 0" "1 8 1" "1 16 1" "1 4 1"; do
    set -- $vv
    v=$1; sms=$2; win=$3
    port=$((port + 1))
    echo "=== FHB_EARLY_REDUCE=$v FHB_COMM_SMS=$sms FHB_COMM_WINDOW=$win rep $rep" >> $LOG
    FHB_EARLY_REDUCE=$v FHB_COMM_SMS=$sms FHB_COMM_WINDOW=$win timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_v${v}_s${sms}_w${win}_$rep.json 2> gpurun_out/${TAG}_err.log
    python - >> $LOG 2>&1 <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_v${v}_s${sms}_w${win}_$rep.json").read().strip().splitlines()[-1])
print("ms_per_step %.3f value %.0f e2e %.0f comm_exposed_ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d.get("comm_exposed_ms", -1)))
PY
  done
done
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_n1.json 2>> gpurun_out/${TAG}_err.log
echo "=== N=1" >> $LOG; python tools/print_bench.py gpurun_out/${TAG}_n1.json >> $LOG 2>&1
cat $LOG
