#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== fused two-step kernel: parity"
timeout 900 python -m pytest tests/test_gpu_fast_parity.py -x -q -k "fused" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_fast_parity.py tests/test_group.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
echo "== speed: 4096^2"
python tools/sweep2d.py o1 "SHLL_FUSE2=0,1"
python tools/sweep2d.py o1 "SHLL_FUSE2=1" "SHLL_ROWS_PER_CHUNK=12,16,20,24,28,36,44"
python tools/sweep2d.py o1 "SHLL_FUSE2=1" "SHLL_TMA_STAGES=2,3" "SHLL_EARLY=1,0"
} 2>&1 | tee $O/r2_20.log
