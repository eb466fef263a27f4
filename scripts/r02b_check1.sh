#!/bin/bash
# round-2 stream-K bring-up: kernel tests, microbench A/B (whole tiles / heuristic / forced), step A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -x --durations=5 > gpurun_out/pytest_kernels.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_kernels.log
for m in 1 0 2; do
  STREAMK=$m CASES="${SKCASES:-L3,L2}" timeout 300 python scripts/bench_igemm.py sk$m > gpurun_out/igemm_sk$m.log 2>&1; echo "igemm sk$m rc=$?"
done
paste -d'\n' gpurun_out/igemm_sk1.log gpurun_out/igemm_sk0.log gpurun_out/igemm_sk2.log | cut -c1-150
bash scripts/ab_step.sh "CTRLV_SPLITK=1" "CTRLV_SPLITK=0" 2>&1 | tee gpurun_out/ab_step.log
