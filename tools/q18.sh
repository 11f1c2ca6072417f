#!/bin/bash
# 2-GPU check: group API on real devices, multi-process suite, and both multi-GPU bench shapes side by side
nvidia-smi -L
python -m pytest tests/test_group.py tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -4
show() { python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); o=d.get('other_mode') or {}
    print('$1', 'Gcu/s=%.1f'%(d['value']/1e9), 'ms/step=%.4f'%d['ms_per_step'], 'e2e=%.1f'%(d['e2e']['value']/1e9), 'strict=%.1f'%(o.get('value',0)/1e9), d['config']['parallelism'])
except Exception as e: print('$1 FAILED', e)"; }
for wl in 2d_o1 2d_o2 1d_o2; do
  python bench.py --workload $wl --steps 300 --warmup 20 --no-cpu-baseline 2>gpurun_out/q18_$wl.1.err | tee gpurun_out/q18_${wl}_1.json | show "$wl N=1"
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --workload $wl --steps 300 --warmup 20 2>gpurun_out/q18_$wl.t.err | tee gpurun_out/q18_${wl}_2_torchrun.json | show "$wl N=2 torchrun"
  python bench.py --gpus 2 --workload $wl --steps 300 --warmup 20 2>gpurun_out/q18_$wl.g.err | tee gpurun_out/q18_${wl}_2_group.json | show "$wl N=2 group"
done
tail -3 gpurun_out/q18_*.err | tail -20
