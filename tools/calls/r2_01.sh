#!/bin/bash
# round 2, GPU call 1: new parity tests + smoke + the new bench line (driver-style K=20 and default)
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2_01_gpu.txt 2>&1
nproc >> $O/r2_01_gpu.txt; free -g | head -2 >> $O/r2_01_gpu.txt
echo "== pytest -m gpu"; 
timeout 1200 python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" > $O/r2_01_pytest.log; tail -15 $O/r2_01_pytest.log; grep "fast long run" $O/r2_01_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12 | tee $O/r2_01_smoke.log
echo "== bench K=20"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_01_bench_k20.json 2> $O/r2_01_bench_k20.err; tail -c 3000 $O/r2_01_bench_k20.json; tail -5 $O/r2_01_bench_k20.err
