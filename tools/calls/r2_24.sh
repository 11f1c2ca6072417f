#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== 1D two-step kernel: parity"
timeout 900 python -m pytest tests/test_gpu_fast_parity.py -x -q -k "fused_two_step_1d" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_parity.py tests/test_group.py -x -q 2>&1 | tail -4
echo "== speed"
python tools/sweep1d.py 67108864 fast "SHLL_FUSE1D=0,1"
python tools/sweep1d.py 8388608 fast "SHLL_FUSE1D=0,1"
python tools/sweep1d.py 67108864 fast "SHLL_FUSE1D=1" "SHLL_1D_TILES_PER_WARP=4,8,16"
} 2>&1 | tee $O/r2_24.log
