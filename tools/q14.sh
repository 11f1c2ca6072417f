python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -3
for tpw in 1 2 4 8 16; do
  for mode in fast strict; do
    SHLL_1D_TILES_PER_WARP=$tpw python bench.py --workload 1d_o1 --mode $mode --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o1 $mode tpw$tpw', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
for tpw in 2 4 16; do
    SHLL_1D_TILES_PER_WARP=$tpw python bench.py --workload 1d_o2 --mode strict --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-other-mode 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2 strict tpw$tpw', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --workload 1d_o2 --steps 200 --warmup 10 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2 N=2', 'Gcu/s=%.1f'%(d['value']/1e9), 'strict=%.1f'%(d['other_mode']['value']/1e9))"
