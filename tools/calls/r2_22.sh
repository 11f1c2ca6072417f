#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu (full)"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "== bench K=20"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_22_bench_k20.json 2> $O/r2_22_bench_k20.err; tail -3 $O/r2_22_bench_k20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_22_bench_k20.json'))
print('headline Gcu/s %.1f frac %.3f launches %s e2e %.1f other %.1f' % (d['value']/1e9, d['roofline']['frac'], d['gpu_launches'], d['e2e']['value']/1e9, d['other_mode']['value']/1e9), d['config']['kernel'], d['clocks'])
for n,b in d['workloads'].items(): print(n, '%.1f Gcu/s frac %.3f' % (b['value']/1e9, b['roofline']['frac']))
PY
