#!/bin/bash
# long-run vs short-run step time of the default workload: developed flow or elapsed time?
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload 2d_o1 --mode fast --steps $1 --warmup $2 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o1 $label steps $1 warmup $2', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], d['clocks'])"; }
b "pdl1" X=1 -- 300 6
b "pdl1" X=1 -- 300 3000
b "pdl1" X=1 -- 3277 20
b "pdl0" SHLL_PDL=0 -- 3277 20
