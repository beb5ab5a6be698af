#!/bin/bash
# compute-sanitizer passes over the small parity tests (the reference has no race / memory checking of its own).
# memcheck: the reference-fixture model test (every kernel of a training step, tiny geometry) + the attention-map kernels;
# racecheck (shared-memory hazards): the attention-map kernels and the LayerNorm / loss kernels.
# usage: tools/gpu_sanitize.sh TAG
TAG=${1:-sanitize}
mkdir -p gpurun_out
SEL_MEM='attention_map_kernels_vs_torch or (model_matches_reference_fixture and nopad) or attention_map_recipe_matches_reference_fixture'
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL_MEM" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/${TAG}_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_memcheck.log | tail -5
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "attention_map_kernels_vs_torch or layernorm_fwd_bwd or distill_loss_and_adamw" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/${TAG}_racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|exit" gpurun_out/${TAG}_racecheck.log | tail -5
