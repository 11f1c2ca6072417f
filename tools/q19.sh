#!/bin/bash
# pipelined multi-step 2D launch: parity first (bounded), then the speed of both launch shapes
export SHLL_HALO_TIMEOUT_MS=2000
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined" 2>&1 | tail -15
b() { local label=$1; shift; local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 python bench.py --workload $1 --mode fast --steps $2 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 $label steps $2', d['config']['kernel'][-30:], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'launches', d['gpu_launches'])
except Exception as e: print('$1 $label FAILED', e)"; }
b pipe0 SHLL_PIPE2D=0 -- 2d_o1 300
b pipe1 SHLL_PIPE2D=1 -- 2d_o1 300
b pipe1 SHLL_PIPE2D=1 -- 2d_o1 3277
b "pipe1 rpc24" SHLL_PIPE2D=1 SHLL_ROWS_PER_CHUNK=24 -- 2d_o1 300
b "pipe1 rpc36" SHLL_PIPE2D=1 SHLL_ROWS_PER_CHUNK=36 -- 2d_o1 300
b "pipe1 stages3" SHLL_PIPE2D=1 SHLL_TMA_STAGES=3 -- 2d_o1 300
b pipe0 SHLL_PIPE2D=0 -- 2d_o2 200
b pipe1 SHLL_PIPE2D=1 -- 2d_o2 200
b "pipe1 rpc128" SHLL_PIPE2D=1 SHLL_ROWS_PER_CHUNK=128 -- 2d_o2 200
