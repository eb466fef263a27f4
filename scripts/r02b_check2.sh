#!/bin/bash
# temporal-attention kernel bring-up: parity tests, microbench A/B against the previous build, step A/B
set -u
mkdir -p gpurun_out
BASE=$PWD/ctrl-v_b200/build/libctrlv_base.so
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x -k "attn" > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -3 gpurun_out/pytest_attn.log
timeout 900 python -m pytest tests/test_gpu_model.py -q -x > gpurun_out/pytest_model.log 2>&1; echo "pytest model rc=$?"; tail -3 gpurun_out/pytest_model.log
echo "--- attn new"; timeout 200 python scripts/bench_attn.py 2>&1 | grep temporal | tee gpurun_out/attn_new.jsonl
echo "--- attn base"; CTRLV_B200_LIB=$BASE timeout 200 python scripts/bench_attn.py 2>&1 | grep temporal | tee gpurun_out/attn_base.jsonl
bash scripts/ab_step.sh "CTRLV_X=new" "CTRLV_B200_LIB=$BASE" 2>&1 | tee gpurun_out/ab_step2.log
