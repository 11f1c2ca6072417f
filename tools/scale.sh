#!/bin/bash
# scaling sweep: bench.py at N = 1, 2, 4, 8 (as many as the box has), for the weak-scaled 2D workloads and the strong-scaled 1D tube
NG=$(nvidia-smi -L | wc -l)
for wl in 2d_o1 2d_o2 1d_o2; do
  for n in 1 2 4 8; do
    [ $n -gt $NG ] && continue
    if [ $n -eq 1 ]; then
      python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline 2>gpurun_out/scale_${wl}_$n.err | tail -1 > gpurun_out/scale_${wl}_$n.json
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --workload $wl --steps 300 --warmup 20 2>gpurun_out/scale_${wl}_$n.err | tail -1 > gpurun_out/scale_${wl}_$n.json
    fi
    python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/scale_${wl}_$n.json')); o=d.get('other_mode') or {}
    print('$wl N=$n', 'Gcu/s=%.1f'%(d['value']/1e9), 'per-GPU=%.1f'%(d['value']/1e9/$n), 'ms/step=%.4f'%d['ms_per_step'], 'e2e=%.1f'%(d['e2e']['value']/1e9), 'strict=%.1f'%(o.get('value',0)/1e9))
except Exception as e:
    print('$wl N=$n FAILED', e); print(open('gpurun_out/scale_${wl}_$n.err').read()[-1500:])
"
  done
done
