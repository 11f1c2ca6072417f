#!/bin/bash
# STRICT with one range guard per cell: the bit-exact suite is the gate, then the STRICT rates
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_host_programs.py -x -q -m gpu 2>&1 | tail -3
for wl in 1d_o1 2d_o1 1d_o2 2d_o2; do
  timeout 200 python bench.py --workload $wl --mode strict --steps 200 --warmup 6 --no-cpu-baseline --no-e2e --no-other-mode 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl strict Gcu/s=%.1f frac=%.3f'%(d['value']/1e9, d['roofline']['frac']))"
done
