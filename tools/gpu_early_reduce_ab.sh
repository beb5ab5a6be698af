#!/bin/bash
# N-GPU A/B of FHB_EARLY_REDUCE (all-reduce of the layer / head gradients under the front-end backward).  usage: N TAG
N=${1:-2}; TAG=${2:-early}
NG=$N
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_ab.log
: > $LOG
port=29520
for rep in 1 2; do
  for vv in "0 6" "0 1" "1 6" "1 2"; do
    set -- $vv
    v=$1; nbk=$2
    port=$((port + 1))
    echo "=== FHB_EARLY_REDUCE=$v FHB_REDUCE_BUCKETS=$nbk rep $rep" >> $LOG
    FHB_EARLY_REDUCE=$v FHB_REDUCE_BUCKETS=$nbk timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_v${v}_b${nbk}_$rep.json 2> gpurun_out/${TAG}_err.log
    python tools/print_bench.py gpurun_out/${TAG}_v${v}_b${nbk}_$rep.json >> $LOG 2>&1
    tail -n 2 gpurun_out/${TAG}_err.log | cut -c1-300 >> $LOG
  done
done
cat $LOG
