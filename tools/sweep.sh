#!/bin/bash
# usage: sweep.sh workload mode "vecs" "rpcs" "stages"
wl=$1; mode=$2
for v in $3; do for rc in $4; do for st in $5; do
SHLL_VEC=$v SHLL_ROWS_PER_CHUNK=$rc SHLL_TMA_STAGES=$st python bench.py --workload $wl --mode $mode --steps 100 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$wl $mode vec $v rpc $rc stages $st', 'Gcu/s=%.1f'%(d['value']/1e9))
except Exception as e: print('$wl $mode vec $v rpc $rc stages $st FAILED')"
done; done; done
