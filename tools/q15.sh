python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for wl in 1d_o1 1d_o2; do
  for mode in strict fast; do
    python bench.py --workload $wl --mode $mode --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl $mode', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
env SHLL_PERSIST=1 python bench.py --workload 1d_o2_64k --steps 20000 --warmup 200 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2_64k persist', 'Gcu/s=%.1f'%(d['value']/1e9), 'us/step=%.3f'%(d['ms_per_step']*1e3), 'launches', d['gpu_launches'], 'strict', d['other_mode']['ms_per_step']*1e3)"
