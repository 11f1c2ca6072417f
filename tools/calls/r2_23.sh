#!/bin/bash
O=gpurun_out; mkdir -p $O
name=2d_o1_acc2_fast
timeout 400 ncu --set full --clock-control none --import-source on -k regex:step2d_acc2 -s 6 -c 1 -o $O/r02_$name -f \
  python bench.py --workload 2d_o1 --mode fast --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-other-mode --no-workloads > $O/r02_ncu_$name.log 2>&1
python tools/ncu_summary.py $O/r02_$name.ncu-rep 33554432 > $O/r02_$name.ncu.txt 2>&1
head -30 $O/r02_$name.ncu.txt
rm -f $O/r02_$name.ncu-rep
echo "== launch list of the default bench command (short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench_default.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_launches_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_bench_default.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    name=r[4].split('(')[0][-60:]
    try: agg[name].append(float(r[-1]))
    except: pass
with open('gpurun_out/r02_launches_bench_default.summary.txt','w') as f:
    for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
        line=f"{k:62s} launches {len(v):5d}  total {sum(v)/1e3:10.1f} us  mean {sum(v)/len(v)/1e3:9.2f} us"
        print(line); f.write(line+"\n")
PY
