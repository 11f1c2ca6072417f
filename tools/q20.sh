#!/bin/bash
# pipelined multi-step 2D launch: parity (bounded), then small / medium grids where launch overhead and ramp/tail dominate
export SHLL_HALO_TIMEOUT_MS=2000
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined" 2>&1 | tail -5
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload $1 --nx $2 --ny $2 --mode fast --steps $3 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 $2^2 $label', 'us/step=%.2f'%(d['ms_per_step']*1e3), 'Gcu/s=%.1f'%(d['value']/1e9), 'launches', d['gpu_launches'], d['config']['kernel'][-34:])
except Exception as e: print('$1 $2 $label FAILED', e)"; }
for wl in 2d_o1 2d_o2; do
  for n in 256 512 1024 2048; do
    b "graph " SHLL_PIPE2D=0 -- $wl $n 2048
    b "pipe  " SHLL_PIPE2D=1 -- $wl $n 2048
  done
done
b "pipe  " SHLL_PIPE2D=1 -- 2d_o1 4096 300
b "pipe rpc36" SHLL_PIPE2D=1 SHLL_ROWS_PER_CHUNK=36 -- 2d_o1 4096 300
