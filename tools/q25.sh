#!/bin/bash
# cache hints on the 2D FAST kernel (results unaffected): L2 promotion of the tensor map, evict-first TMA loads, streaming stores
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload $1 --mode fast --steps 300 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 $label', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'])
except Exception as e: print('$1 $label FAILED', e)"; }
b "default" X=1 -- 2d_o1
b "promo none" SHLL_TMA_L2PROMO=0 -- 2d_o1
b "promo 64" SHLL_TMA_L2PROMO=1 -- 2d_o1
b "promo 256" SHLL_TMA_L2PROMO=3 -- 2d_o1
b "evict loads" SHLL_EVICT=1 -- 2d_o1
b "cs stores" SHLL_EVICT=2 -- 2d_o1
b "evict loads + cs stores" SHLL_EVICT=3 -- 2d_o1
b "promo 256 + evict 3" SHLL_TMA_L2PROMO=3 SHLL_EVICT=3 -- 2d_o1
b "default" X=1 -- 2d_o2
b "promo 256" SHLL_TMA_L2PROMO=3 -- 2d_o2
b "evict loads + cs stores" SHLL_EVICT=3 -- 2d_o2
