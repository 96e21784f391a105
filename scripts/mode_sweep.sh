#!/bin/bash
# timing of the fast-kernel variants on the C5 workload (2^22 candidates x 2 GPs per step); experiments only
B="python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-e2e --log2m ${LOG2M:-22}"
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 160 $B 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('   ms_per_step', round(j['ms_per_step'],3), 'best', j['best']['value'], 'clocks', j['clocks'].get('sm_mhz'), 'W', j['clocks'].get('power_w_max'))
"
done
