#!/bin/bash
# persistent 2D acc kernel: parity, then persistent vs one-block-per-item, rows per chunk, ring depth
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
b() { # label env... -- workload
  local label=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --workload $1 --mode fast --steps 150 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$1 $label', d['config']['kernel'][-24:], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'])
except Exception as e: print('$1 $label FAILED', e)"
}
b persist0 SHLL_PERSIST2D=0 -- 2d_o1
b persist1 SHLL_PERSIST2D=1 -- 2d_o1
for rpc in 12 24 36 64; do b "persist1 rpc$rpc" SHLL_ROWS_PER_CHUNK=$rpc -- 2d_o1; done
b "persist1 stages3 rpc18" SHLL_TMA_STAGES=3 -- 2d_o1
b "persist1 stages3 rpc36" SHLL_TMA_STAGES=3 SHLL_ROWS_PER_CHUNK=36 -- 2d_o1
b persist0 SHLL_PERSIST2D=0 -- 2d_o2
b persist1 SHLL_PERSIST2D=1 -- 2d_o2
for rpc in 32 128 256; do b "persist1 rpc$rpc" SHLL_ROWS_PER_CHUNK=$rpc -- 2d_o2; done
b "persist1 stages3" SHLL_TMA_STAGES=3 -- 2d_o2
b "persist1 cfg0" SHLL_ACC_CFG=0 -- 2d_o2
