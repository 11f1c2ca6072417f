python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
for wl in 2d_o1 2d_o2; do
    python bench.py --workload $wl --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl fast', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
for cfg in 0 3; do
    SHLL_ACC_CFG=$cfg python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg$cfg', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
for rpc in 32 48 96; do
    SHLL_ROWS_PER_CHUNK=$rpc python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast rpc$rpc', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
