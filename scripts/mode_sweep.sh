#!/bin/bash
# timing of the fast-kernel variants on the C5 workload (2^22 candidates x 2 GPs per step); experiments only
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --log2m 22"
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg timeout 120 $B 2>&1 | python -c "import sys,json; [print('   ms_per_step', json.loads(l)['ms_per_step'], 'best', json.loads(l)['best']) for l in sys.stdin if l.startswith('{')]"
done
