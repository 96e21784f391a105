#!/bin/bash
# f8c kernel experiment: parity on ragged shapes, then bench lines for the two generator-warp counts and the
# K1-only / MMA-only timing modes.  Usage (on the GPU box): bash scripts/exp_fast8.sh TAG
TAG=${1:-exp}
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-e2e --no-extras"
mkdir -p gpurun_out
for gw in 16 8; do
  OMBO_FAST_GEN_WARPS=$gw timeout 100 python scripts/f8c_check.py > gpurun_out/${TAG}_check$gw.log 2>&1 || echo "CHECK gw$gw FAILED"
  tail -3 gpurun_out/${TAG}_check$gw.log
  OMBO_FAST_GEN_WARPS=$gw timeout 90 $B > gpurun_out/${TAG}_bench_gw$gw.json 2> gpurun_out/${TAG}_bench_gw$gw.err
  for dbg in 1 2 ${EXTRA_DBG}; do
    OMBO_FAST_GEN_WARPS=$gw OMBO_FAST_DBG=$dbg timeout 90 $B > gpurun_out/${TAG}_bench_gw${gw}_dbg$dbg.json 2> gpurun_out/${TAG}_bench_gw${gw}_dbg$dbg.err
  done
  OMBO_FAST_GEN_WARPS=$gw OMBO_FAST_PROFILE=1 timeout 90 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fit --no-e2e --no-extras --log2m 20 > gpurun_out/${TAG}_prof_gw$gw.json 2> gpurun_out/${TAG}_prof_gw$gw.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"], round(d["roofline"]["avg_launch_ms"],3))
    except Exception as e: print(f, "ERR", e)
PY
tail -2 gpurun_out/${TAG}_prof_gw16.err gpurun_out/${TAG}_prof_gw8.err
