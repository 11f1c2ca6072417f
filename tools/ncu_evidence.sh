#!/bin/bash
# usage: gpurun -- tools/ncu_evidence.sh    round-2 ncu evidence: one --set full capture per hot kernel + the launch list of the default bench command
O=gpurun_out; mkdir -p $O
prof() {  # name regex workload mode cells
  local name=$1 rx=$2 wl=$3 mode=$4 cells=$5
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 6 -c 1 -o $O/r02_$name -f \
    python bench.py --workload $wl --mode $mode --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-other-mode --no-workloads > $O/r02_ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/r02_$name.ncu-rep $cells > $O/r02_$name.ncu.txt 2>&1
  head -9 $O/r02_$name.ncu.txt | tail -8
  ncu -i $O/r02_$name.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_$name.source.csv.gz
  rm -f $O/r02_$name.ncu-rep
}
prof 2d_o1_acc2_fast 'step2d_acc2' 2d_o1 fast 33554432      # two steps per launch: cells x 2
SHLL_FUSE2=0 prof 2d_o1_acc_fast 'step2d_acc' 2d_o1 fast 16777216
prof 2d_o2_acc_fast 'step2d_acc' 2d_o2 fast 33554432
prof 2d_o1_tma_strict 'step2d_tma' 2d_o1 strict 16777216
prof 1d_o2_acc2_fast 'step1d_acc2' 1d_o2 fast 134217728
SHLL_FUSE1D=0 prof 1d_o2_acc_fast 'step1d' 1d_o2 fast 67108864
echo "== launch list of the default bench command (short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench_default.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_bench_default.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    name=r[4].split('(')[0][-60:]; 
    try: agg[name].append(float(r[-1]))
    except: pass
with open('gpurun_out/r02_launches_bench_default.summary.txt','w') as f:
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        line=f"{k:62s} launches {len(v):5d}  total {sum(v)/1e3:10.1f} us  mean {sum(v)/len(v)/1e3:9.2f} us"
        print(line); f.write(line+"\n")
PY
du -sm $O | cut -f1
