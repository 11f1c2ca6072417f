python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or variant or ragged" 2>&1 | tail -5
for wl in 2d_o1 2d_o2; do
  for vec in 1 2; do
    SHLL_VEC=$vec python bench.py --workload $wl --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl fast vec$vec', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
