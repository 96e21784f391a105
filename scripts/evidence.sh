#!/bin/bash
# Round evidence in one GPU call: GPU tests, the default bench line, the ncu launch list of a short bench run and one
# `ncu --set full` capture of the headline kernel.  Usage (on the GPU box): bash scripts/evidence.sh TAG [kernel-regex]
TAG=${1:-r02}
KREGEX=${2:-k_posterior_fast8}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-fit --no-e2e --no-extras --no-refresh-timing"
timeout 300 $SHORT > gpurun_out/${TAG}_short.json 2> gpurun_out/${TAG}_short.err; echo "short rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s 6 -c 1 -f -o gpurun_out/${TAG}_fast8 $SHORT > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("ms", round(d["ms_per_step"], 2), "frac", round(d["roofline"]["frac"], 4), "clk", d["clocks"], "e2e", d["e2e"]["value"] if d.get("e2e") else None)
PY
