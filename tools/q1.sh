python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or variant or ragged or geometry" 2>&1 | tail -15
for wl in 2d_o1 2d_o2; do
  for acc in 0 1; do
    SHLL_VEC=2 SHLL_ACC=$acc python bench.py --workload $wl --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl fast acc$acc', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
