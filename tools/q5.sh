python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for cfg in 0 1 2 3 4; do
    SHLL_ACC_CFG=$cfg python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg$cfg', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
for st in 2 3 4 5; do
    SHLL_TMA_STAGES=$st python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg1 stages$st', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
for rpc in 32 128 256; do
    SHLL_ROWS_PER_CHUNK=$rpc python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg1 rpc$rpc', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
python bench.py --workload 2d_o1 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o1 fast', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
