for cfg in 0 1 2; do
  for tpw in 4 8; do
    SHLL_ACC_CFG=$cfg SHLL_1D_TILES_PER_WARP=$tpw python bench.py --workload 1d_o2 --mode fast --steps 200 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('1d_o2 fast cfg$cfg tpw$tpw', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
  done
done
ncu --set full --clock-control none --import-source on -k regex:step1d_acc -s 3 -c 1 -o gpurun_out/prof_1d_o2_acc_r8 -f python bench.py --workload 1d_o2 --mode fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_acc1d.log 2>&1
tail -1 gpurun_out/ncu_acc1d.log
