#!/bin/bash
# quick GPU iteration: parity tests + dynamic instruction counts + timings for the main kernels
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "strict or random or ragged or variant or geometry or general or denormal" 2>&1 | tail -12
for wl in 2d_o1 2d_o2 1d_o1 1d_o2; do
  for mode in strict fast; do
    python bench.py --workload $wl --mode $mode --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl $mode', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
for v in "SHLL_PERSIST=0 SHLL_GRAPH=0" "SHLL_PERSIST=0 SHLL_GRAPH=1" "SHLL_PERSIST=1"; do
  env $v python bench.py --workload 1d_o2_64k --steps 20000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2_64k [$v]', 'Gcu/s=%.1f'%(d['value']/1e9), 'us/step=%.3f'%(d['ms_per_step']*1e3), 'launches', d['gpu_launches'])"
done
