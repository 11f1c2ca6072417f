#!/bin/bash
# usage: gpurun --gpus 8 -- tools/scale8.sh    multi-GPU parity (2, 4, 8 GPUs) + the driver-style bench line at N = 8 and N = 1 on ONE box -> gpurun_out/r02_scale_n{1,8}.json
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3 | tee $O/r2_25_pytest_mgpu.log
N=8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_scale_n$N.json 2> $O/r02_scale_n$N.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_scale_n1.json 2> $O/r02_scale_n1.err
python - <<'PY'
import json
base={}
for N in (1,8):
    try:
        d=[json.loads(l) for l in open(f'gpurun_out/r02_scale_n{N}.json') if l.startswith('{')][0]
    except Exception as e:
        print(N, 'failed', e); print(open(f'gpurun_out/r02_scale_n{N}.err').read()[-1500:]); continue
    row={'2d_o1': d['value'], **{k:v['value'] for k,v in d['workloads'].items() if 'value' in v}}
    if N==1: base=row
    print(f"N={N}: " + "  ".join(f"{k} {v/1e9:8.1f} G eff {v/(N*base.get(k,v/N)):.3f}" for k,v in row.items()),
          '| parity', d.get('parity_vs_1gpu'), {k:v.get('parity_vs_1gpu') for k,v in d['workloads'].items()}, '| e2e %.1f G' % (d['e2e']['value']/1e9))
PY
