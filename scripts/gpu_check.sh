#!/bin/bash
# GPU check: the -m gpu suite, smoke() under an ncu launch list (the first launches must be this
# library's kernels), and every bench.py mode.  Run from the repo root on a B200 box.
#   FAST=1: tests + config-2 bench only
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json
if [ "${FAST:-0}" = "1" ]; then exit 0; fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/smoke_launches.csv \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
python bench.py --config 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 rc=$?"; cat gpurun_out/bench_c3.json
python bench.py --config 4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; cat gpurun_out/bench_c4.json
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/bench_c5.jsonl 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; cut -c1-200 gpurun_out/bench_c5.jsonl
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
