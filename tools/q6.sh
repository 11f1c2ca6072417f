for rpc in 18 22 26 30 38 46; do
    SHLL_ROWS_PER_CHUNK=$rpc python bench.py --workload 2d_o1 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o1 fast acc rpc$rpc', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
for st in 2 4; do
    SHLL_ROWS_PER_CHUNK=26 SHLL_TMA_STAGES=$st python bench.py --workload 2d_o1 --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('2d_o1 fast acc rpc26 stages$st', 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:step2d_acc -s 3 -c 1 -o gpurun_out/prof_2d_o2_acc_r7 -f python bench.py --workload 2d_o2 --mode fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step2d_acc -s 3 -c 1 -o gpurun_out/prof_2d_o1_acc_r7 -f python bench.py --workload 2d_o1 --mode fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e >> gpurun_out/ncu_acc.log 2>&1
tail -2 gpurun_out/ncu_acc.log
