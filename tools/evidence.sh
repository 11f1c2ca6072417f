#!/bin/bash
# One GPU call that refreshes everything the round is judged on: the GPU test suite, the default bench line and the
# reference arm, every workload in both arithmetic modes, the ncu launch list of the default bench command and one
# `ncu --set full` capture per hot kernel (summarised afterwards with tools/ncu_summary.py into profiles/).
# usage (from the repo root, on a GPU box): tools/evidence.sh [tag]       -> files under gpurun_out/<tag>_*
TAG=${1:-ev}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -x -q -m gpu > $O/${TAG}_pytest_gpu.log 2>&1; tail -3 $O/${TAG}_pytest_gpu.log
echo "== bench default"
timeout 600 python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; tail -c 1500 $O/${TAG}_bench_default.json
echo "== reference arm"
timeout 600 python bench.py --impl reference > $O/${TAG}_bench_reference.json 2>&1; tail -c 600 $O/${TAG}_bench_reference.json
echo "== all workloads"
: > $O/${TAG}_all_workloads.log
for wl in 2d_o1 2d_o2 1d_o1 1d_o2; do
  timeout 300 python bench.py --workload $wl --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 > $O/${TAG}_wl_$wl.json
  python - <<EOF | tee -a $O/${TAG}_all_workloads.log
import json
d=json.load(open('$O/${TAG}_wl_$wl.json')); o=d['other_mode']
print('$wl fast  ', d['config']['kernel'], 'Gcu/s=%.1f'%(d['value']/1e9), 'frac=%.3f'%d['roofline']['frac'], 'ms=%.4f'%d['ms_per_step'], 'e2e=%.1f'%(d['e2e']['value']/1e9))
print('$wl strict', o['kernel'], 'Gcu/s=%.1f'%(o['value']/1e9), 'frac=%.3f'%o['roofline_frac'], 'ms=%.4f'%o['ms_per_step'])
EOF
done
for v in "SHLL_PERSIST=0 SHLL_GRAPH=0" "SHLL_PERSIST=0 SHLL_GRAPH=1" "SHLL_PERSIST=1"; do
  env $v timeout 300 python bench.py --workload 1d_o2_64k --steps 20000 --warmup 200 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); o=d['other_mode']; print('1d_o2_64k [$v] fast us/step=%.3f'%(d['ms_per_step']*1e3), 'Gcu/s=%.1f'%(d['value']/1e9), 'launches', d['gpu_launches'], '| strict us/step=%.3f'%(o['ms_per_step']*1e3), 'Gcu/s=%.1f'%(o['value']/1e9))" | tee -a $O/${TAG}_all_workloads.log
done
echo "== ncu launch list of the default bench command (short run)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_bench_default.csv \
  python bench.py --steps 100 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches_bench.log 2>&1
echo "== ncu --set full, one launch per hot kernel"
prof() {  # name regex workload mode [env...]
  local name=$1 rx=$2 wl=$3 mode=$4; shift 4
  env "$@" timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s 3 -c 1 -o $O/${TAG}_$name -f \
    python bench.py --workload $wl --mode $mode --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode > $O/${TAG}_ncu_$name.log 2>&1
  ls -la $O/${TAG}_$name.ncu-rep 2>&1 | tail -1
  python tools/ncu_summary.py $O/${TAG}_$name.ncu-rep $CELLS > $O/${TAG}_$name.ncu.txt 2>&1
  head -8 $O/${TAG}_$name.ncu.txt
}
WANT=${EVIDENCE_PROFILES:-"2d_o1_fast 2d_o1_strict 2d_o2_fast 2d_o2_strict 1d_o2_fast 1d_o1_strict"}
for name in $WANT; do
  case $name in
    2d_o1_*) CELLS=16777216; rx='step2d_(acc|tma)'; wl=2d_o1 ;;
    2d_o2_*) CELLS=33554432; rx='step2d_(acc|tma)'; wl=2d_o2 ;;
    1d_o2_*) CELLS=67108864; rx='step1d'; wl=1d_o2 ;;
    1d_o1_*) CELLS=67108864; rx='step1d'; wl=1d_o1 ;;
  esac
  prof $name "$rx" $wl ${name##*_} X=1
done
SZ=$(du -sm $O | cut -f1); echo "gpurun_out is $SZ MiB"
if [ "$SZ" -gt 55 ]; then rm -f $O/${TAG}_*strict.ncu-rep; echo "dropped the strict .ncu-rep files (summaries kept)"; fi
echo "== done"
