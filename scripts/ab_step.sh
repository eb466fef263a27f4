#!/bin/bash
# A/B of the config-2 step under environment switches, interleaved on one box:
#   bash scripts/ab_step.sh "CTRLV_FF_CG=1" "CTRLV_FF_CG=2" ...   -> one line per run (ms/step, SM clock)
for rep in 1 2; do
  for v in "$@"; do
    env $v timeout -s KILL 150 python bench.py --steps 10 --warmup 3 2>/dev/null > /tmp/ab.json
    python - "$v" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/ab.json"))
    print(sys.argv[1], "ms_per_step %.3f" % d["ms_per_step"], "sm_mhz", d["clocks"]["sm_mhz"], "launches", d["launches_per_step"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
