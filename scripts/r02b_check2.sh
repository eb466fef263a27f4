#!/bin/bash
# fused LayerNorm + FeedForward bring-up: parity tests, microbench, step A/B against the previous build
set -u
mkdir -p gpurun_out
BASE=$PWD/ctrl-v_b200/build/libctrlv_base.so
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "feedforward or attn" > gpurun_out/pytest_ff.log 2>&1; echo "pytest ff rc=$?"; tail -3 gpurun_out/pytest_ff.log
timeout 600 python -m pytest tests/test_gpu_model.py -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/pytest_model.log
for rep in 1 2; do for v in "CTRLV_FF_LN=1" "CTRLV_FF_LN=0"; do
  env $v timeout -s KILL 150 python bench.py --steps 10 --warmup 3 2>/dev/null > /tmp/ab.json
  python -c "import json,sys; d=json.load(open('/tmp/ab.json')); print(sys.argv[1], 'ms_per_step %.3f'%d['ms_per_step'], 'sm_mhz', d['clocks']['sm_mhz'], 'launches', d['launches_per_step'])" "$v" || echo "$v FAILED"
done; done 2>&1 | tee gpurun_out/ab_step3.log
