#!/bin/bash
# round-end rehearsal: what the driver runs on a fresh box
nproc; free -g | head -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
(time python bench.py --impl reference) > gpurun_out/q17_ref.json 2> gpurun_out/q17_ref.err; tail -c 1200 gpurun_out/q17_ref.json; tail -4 gpurun_out/q17_ref.err
(time python bench.py) > gpurun_out/q17_bench.json 2> gpurun_out/q17_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/q17_bench.json').read().strip().splitlines()[-1])
print('value %.1f G  e2e %.1f G  frac %.3f  launches %d' % (d['value']/1e9, d['e2e']['value']/1e9, d['roofline']['frac'], d['gpu_launches']))
print('clocks', d['clocks']); print('cpu_baseline', json.dumps(d['cpu_baseline'])[:900])
PY
tail -4 gpurun_out/q17_bench.err
