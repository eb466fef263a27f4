#!/bin/bash
# ncu evidence of the current build (run from the repo root on a B200 box; results -> gpurun_out/):
#   1. launch list of one eager step with DRAM bytes per launch (-> scripts/dram_traffic.py, summarize_launches.py)
#   2. `--set full` captures of the kernels the roofline statements are about (scripts/ncu_all.sh)
set -u
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/step_launches_dram.csv python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list rc=$?"; grep -c igemm gpurun_out/step_launches_dram.csv
bash scripts/ncu_all.sh conv ff res320 attn tattn norms > gpurun_out/ncu_all.log 2>&1; tail -8 gpurun_out/ncu_all.log
