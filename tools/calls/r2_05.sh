#!/bin/bash
# 8 GPUs: the driver-style scaling line (headline + workloads block + parity) and the 1D strong-scaling detail
O=gpurun_out; mkdir -p $O
N=${1:-8}
nvidia-smi -L | head -8 > $O/r2_05_gpus.txt; nproc >> $O/r2_05_gpus.txt; free -g | head -2 >> $O/r2_05_gpus.txt
echo "== bench --gpus $N (driver style)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_05_bench_n$N.json 2> $O/r2_05_bench_n$N.err
tail -c 1200 $O/r2_05_bench_n$N.err
python - <<PY
import json
try:
    d=[json.loads(l) for l in open('$O/r2_05_bench_n$N.json') if l.startswith('{')][0]
    print('headline', d['config']['workload'], 'Gcu/s', d['value']/1e9, 'frac', d['roofline']['frac'], 'parity', d.get('parity_vs_1gpu'), 'e2e', d['e2e']['value']/1e9, d['e2e'].get('breakdown_s_max_over_ranks'))
    for n,b in d['workloads'].items(): print(n, {k:b.get(k) for k in ('value','ms_per_step','parity_vs_1gpu','reps','error')}, b.get('roofline',{}).get('frac'))
except Exception as e: print('parse failed', e)
PY
