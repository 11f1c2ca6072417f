#!/bin/bash
O=gpurun_out; mkdir -p $O
echo "== bench default (natural run length)"
( time timeout 900 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err ) 2>&1 | grep real
echo "== reference arm"
( time timeout 900 python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json')); r=json.load(open('gpurun_out/r02_bench_reference.json'))
print('default: steps', d['steps'], 'reps', d['reps'], 'Gcu/s %.1f frac %.3f e2e %.1f clocks %s' % (d['value']/1e9, d['roofline']['frac'], d['e2e']['value']/1e9, d['clocks']))
print('other', d['other_mode']['value']/1e9, 'cpu', d['cpu_baseline']['value']/1e6, d['cpu_baseline']['all_cores']['value']/1e6)
print('reference arm: %.1f Mcu/s cores %s' % (r['value']/1e6, r['cpu_baseline']['cores']))
for n,b in d['workloads'].items(): print(n, '%.1f Gcu/s frac %.3f' % (b['value']/1e9, b['roofline']['frac']), b['clocks']['reasons'])
PY
