#!/bin/bash
# 8 GPUs: 1D strong scaling with / without temporal halo blocking
O=gpurun_out; mkdir -p $O
N=${1:-8}
for K in 16; do
  SHLL_HALO_K=$K timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+K)) bench.py --gpus $N --workload 1d_o2 --steps 100 --warmup 10 --no-e2e --no-other-mode --no-workloads > $O/r2_08_1d_n${N}_k$K.json 2> $O/r2_08_1d_n${N}_k$K.err
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$O/r2_08_1d_n${N}_k$K.json') if l.startswith('{')][0]
    print('K=$K N=$N 1d_o2: Gcu/s %.1f us/step %.2f parity %s' % (d['value']/1e9, d['ms_per_step']*1e3, d.get('parity_vs_1gpu')), d.get('halo_exchange'))
except Exception as e:
    print('parse failed', e); print(open('$O/r2_08_1d_n${N}_k$K.err').read()[-1500:])
PY
done
