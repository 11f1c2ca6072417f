#!/bin/bash
# chunk-height heuristic + PDL: whole GPU suite, then small / medium grids in both modes, PDL inside the graph
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload $1 --nx $2 --ny $3 --mode fast --steps $4 --warmup 6 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); o=d['other_mode']; print('$1 $2x$3 $label', 'fast us/step=%.2f'%(d['ms_per_step']*1e3), 'Gcu/s=%.1f'%(d['value']/1e9), d['config']['kernel'][-22:], '| strict us/step=%.2f'%(o['ms_per_step']*1e3), 'Gcu/s=%.1f'%(o['value']/1e9), o['kernel'][-22:])
except Exception as e: print('$1 $2 $label FAILED', e)"; }
for n in 256 512 1024 2048; do
  b "auto" X=1 -- 2d_o1 $n $n 2048
  b "auto" X=1 -- 2d_o2 $n $n 1024
done
b "auto pdl-in-graph" SHLL_PDL_GRAPH=1 -- 2d_o1 256 256 2048
b "auto pdl-in-graph" SHLL_PDL_GRAPH=1 -- 2d_o1 1024 1024 2048
b "auto pdl-in-graph" SHLL_PDL_GRAPH=1 -- 2d_o2 1024 1024 1024
b "auto nograph" SHLL_GRAPH=0 -- 2d_o1 256 256 2048
b "auto nograph" SHLL_GRAPH=0 -- 2d_o1 1024 1024 2048
b "auto nograph" SHLL_GRAPH=0 -- 2d_o2 1024 1024 1024
b "auto" X=1 -- 2d_o1 4096 4096 400
