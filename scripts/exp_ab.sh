#!/bin/bash
# same-box A/B of bench lines under different environment settings: scripts/exp_ab.sh TAG "ENV1" "ENV2" ...
TAG=$1; shift
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-e2e --no-extras"
i=0
for rep in 1 2; do
  for envs in "$@"; do
    i=$((i+1))
    env $envs timeout 90 $B > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$i.json")); print("$envs", "| ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"], "launch ms", round(d["roofline"]["avg_launch_ms"],3), "launches", d["roofline"]["launches"])
except Exception as e: print("$envs", "ERR", e)
PY
  done
done
