#!/bin/bash
# same-box A/B of library builds: scripts/exp_libs.sh TAG libA.so libB.so ...   (files under scripts/trace_lib/)
TAG=$1; shift
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-e2e --no-extras"
for rep in 1 2; do
  for lib in "$@"; do
    mkdir -p scripts/trace_lib; cp scripts/trace_lib/$lib optimobo_b200/liboptimobo_b200.so
    timeout 90 $B > gpurun_out/${TAG}_${lib}_$rep.json 2> gpurun_out/${TAG}_${lib}_$rep.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${lib}_$rep.json")); print("$lib", "| ms", round(d["ms_per_step"],2), "clk", d["clocks"]["sm_mhz"], "W", d["clocks"]["power_w_max"], "launch ms", round(d["roofline"]["avg_launch_ms"],3))
except Exception as e: print("$lib", "ERR", e)
PY
  done
done
