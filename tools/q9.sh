python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for acc in 0 1; do
    SHLL_ACC=$acc python bench.py --workload 1d_o2 --mode fast --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2 fast acc$acc', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
bash tools/q3.sh 2>&1 | grep -A12 "ACC=1" | grep "1d_o2"
