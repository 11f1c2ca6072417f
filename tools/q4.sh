bash tools/q3.sh 2>&1 | grep -A20 "SHLL_ACC=1" | grep "2d_"
for wl in 2d_o1 2d_o2; do
    python bench.py --workload $wl --mode fast --steps 100 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$wl fast', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'])"
done
ncu --set full --clock-control none --import-source on -k regex:step2d_acc -s 3 -c 1 -o gpurun_out/prof_2d_o2_acc_r6 -f python bench.py --workload 2d_o2 --mode fast --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_acc.log 2>&1
tail -2 gpurun_out/ncu_acc.log
