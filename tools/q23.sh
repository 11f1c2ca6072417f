#!/bin/bash
# PDL on every single-GPU step kernel: whole GPU suite, then every workload in both modes, PDL on / off
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for pdl in 0 1; do
for wl in 2d_o1 2d_o2 1d_o1 1d_o2; do
  SHLL_PDL=$pdl timeout 300 python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_mode']
print('pdl$pdl $wl fast Gcu/s=%.1f frac=%.3f | strict Gcu/s=%.1f frac=%.3f'%(d['value']/1e9, d['roofline']['frac'], o['value']/1e9, o['roofline_frac']))"
done
done
