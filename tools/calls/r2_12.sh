#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_fast_parity.py tests/test_group.py -x -q 2>&1 | tail -3
echo "== 2D o2 FAST with one shuffle level less"
python tools/sweep2d.py o2 "SHLL_EARLY=1,0"
} 2>&1 | tee $O/r2_12.log
