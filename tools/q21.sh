#!/bin/bash
# (1) programmatic dependent launch on / off; (2) chunk height on small and medium 2D grids
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2d or golden or random" 2>&1 | tail -3
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload $1 --nx $2 --ny $3 --mode fast --steps $4 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 $2x$3 $label', 'us/step=%.2f'%(d['ms_per_step']*1e3), 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], d['config']['kernel'][-22:])
except Exception as e: print('$1 $2 $label FAILED', e)"; }
b "pdl0" SHLL_PDL=0 -- 2d_o1 4096 4096 400
b "pdl1" SHLL_PDL=1 -- 2d_o1 4096 4096 400
b "pdl0" SHLL_PDL=0 -- 2d_o1 4096 4096 400
b "pdl1" SHLL_PDL=1 -- 2d_o1 4096 4096 400
b "pdl0" SHLL_PDL=0 -- 2d_o2 2048 16384 200
b "pdl1" SHLL_PDL=1 -- 2d_o2 2048 16384 200
for n in 256 1024 2048; do
  for rpc in 2 6 10 18; do b "o1 rpc$rpc" SHLL_ROWS_PER_CHUNK=$rpc -- 2d_o1 $n $n 2048; done
  for rpc in 4 8 16 64; do b "o2 rpc$rpc" SHLL_ROWS_PER_CHUNK=$rpc -- 2d_o2 $n $n 1024; done
done
b "o1 rpc18 nograph" SHLL_GRAPH=0 -- 2d_o1 1024 1024 2048
b "o1 rpc6 nograph" SHLL_GRAPH=0 SHLL_ROWS_PER_CHUNK=6 -- 2d_o1 1024 1024 2048
