python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fast or geometry or ragged" 2>&1 | tail -3
for cfg in 1 2 3 4; do
  for st in 2 3; do
    SHLL_TMA_STAGES=$st SHLL_ACC_CFG=$cfg python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg$cfg stages$st', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
for rpc in 10 14 18; do
  for st in 2; do
    SHLL_TMA_STAGES=$st SHLL_ROWS_PER_CHUNK=$rpc python bench.py --workload 2d_o1 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o1 fast acc rpc$rpc stages$st', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
for rpc in 32 48; do
    SHLL_ACC_CFG=3 SHLL_ROWS_PER_CHUNK=$rpc python bench.py --workload 2d_o2 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o2 fast cfg3 rpc$rpc', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
