#!/bin/bash
O=gpurun_out; mkdir -p $O
name=1d_o2_acc2_fast
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step1d_acc2 -s 6 -c 1 -o $O/r02_$name -f \
  python bench.py --workload 1d_o2 --mode fast --steps 12 --warmup 4 --no-cpu-baseline --no-e2e --no-other-mode --no-workloads > $O/r02_ncu_$name.log 2>&1
python tools/ncu_summary.py $O/r02_$name.ncu-rep 134217728 > $O/r02_$name.ncu.txt 2>&1
head -12 $O/r02_$name.ncu.txt; grep "stall reasons\|per 32 cell" $O/r02_$name.ncu.txt | head -3
rm -f $O/r02_$name.ncu-rep
echo "== pytest -m gpu (full, final)"
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/r02_pytest_gpu_final.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee $O/r02_smoke_final.log
