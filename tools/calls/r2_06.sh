#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== 1D temporal halo blocking + group tests (slabs sharing device 0)"
timeout 900 python -m pytest tests/test_group.py -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_parity.py -x -q -k "1d or persistent or graph" 2>&1 | tail -3
echo "== single GPU, 2^23 cells (the per-GPU share at N=8): time per step without any exchange"
python bench.py --workload 1d_o2 --nx 8388608 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-other-mode --no-workloads | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1d_o2 2^23 cells: us/step', d['ms_per_step']*1e3, 'Gcu/s', d['value']/1e9)"
} 2>&1 | tee $O/r2_06.log
