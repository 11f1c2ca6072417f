#!/bin/bash
# full GPU suite + smoke + README-table timings + default bench (driver style and full run)
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"
timeout 1500 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" > $O/r2_14_pytest.log; tail -4 $O/r2_14_pytest.log; grep "fast long run" $O/r2_14_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== README tables"
timeout 900 python tools/readme_tables.py $O/r02_readme_tables.json 2>&1 | tail -20
echo "== bench K=20"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_14_bench_k20.json 2> $O/r2_14_bench_k20.err; tail -3 $O/r2_14_bench_k20.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_14_bench_k20.json'))
print('headline Gcu/s %.1f frac %.3f e2e %.1f other %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, {k:d['other_mode'][k] for k in ('value','roofline_frac')}))
for n,b in d['workloads'].items(): print(n, '%.1f Gcu/s frac %.3f' % (b['value']/1e9, b['roofline']['frac']))
PY
