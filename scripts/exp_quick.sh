#!/bin/bash
# quick A/B of the f8c kernel: parity on ragged shapes + one bench line per generator-warp count
TAG=${1:-q}
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-e2e --no-extras"
for gw in ${GWS:-16 8}; do
  OMBO_FAST_GEN_WARPS=$gw timeout 100 python scripts/f8c_check.py > gpurun_out/${TAG}_check$gw.log 2>&1 || echo "CHECK gw$gw FAILED"
  tail -1 gpurun_out/${TAG}_check$gw.log
  OMBO_FAST_GEN_WARPS=$gw timeout 90 $B > gpurun_out/${TAG}_bench_gw$gw.json 2> gpurun_out/${TAG}_bench_gw$gw.err
  for dbg in ${DBGS}; do
    OMBO_FAST_GEN_WARPS=$gw OMBO_FAST_DBG=$dbg timeout 90 $B > gpurun_out/${TAG}_bench_gw${gw}_dbg$dbg.json 2> gpurun_out/${TAG}_bench_gw${gw}_dbg$dbg.err
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], round(d["roofline"]["avg_launch_ms"],3))
    except Exception as e: print(f, "ERR", e)
PY
