#!/bin/bash
# PDL with slabs: 2-GPU suites, then N=2 with and without it
export SHLL_HALO_TIMEOUT_MS=3000
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_group.py -x -q -m gpu 2>&1 | tail -3
for pdl in 0 1; do
for wl in 1d_o2 2d_o1; do
  SHLL_PDL_MULTI=$pdl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2970$pdl bench.py --gpus 2 --workload $wl --steps 400 --warmup 20 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_mode']
print('N=2 pdl_multi=$pdl $wl fast Gcu/s=%.1f ms/step=%.4f | strict Gcu/s=%.1f'%(d['value']/1e9, d['ms_per_step'], o['value']/1e9))"
done
done
