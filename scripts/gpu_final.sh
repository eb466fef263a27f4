#!/bin/bash
# round-end evidence in one box: the full GPU check, then the ncu launch list of one eager step of the same build
set -u
bash scripts/gpu_check.sh
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/step_launches_dram.csv python scripts/profile_step.py > gpurun_out/profile_step.log 2>&1
echo "launch list rc=$?"; grep -c "gpu__time_duration" gpurun_out/step_launches_dram.csv
