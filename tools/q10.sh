python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "1d or persistent or ragged or 64m" 2>&1 | tail -3
for tpw in 1 2 4 8 16 32; do
    SHLL_1D_TILES_PER_WARP=$tpw python bench.py --workload 1d_o2 --mode fast --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2 fast tpw$tpw', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
