#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== 1D FAST o2, 2^23 cells: tiles per warp"
python tools/sweep1d.py 8388608 fast "SHLL_1D_TILES_PER_WARP=1,2,3,4,6,8,12,16"
echo "== 2^24"
python tools/sweep1d.py 16777216 fast "SHLL_1D_TILES_PER_WARP=2,4,8,16"
echo "== 2^26"
python tools/sweep1d.py 67108864 fast "SHLL_1D_TILES_PER_WARP=4,8,16"
echo "== 2^23 without PDL"
python tools/sweep1d.py 8388608 fast "SHLL_PDL=0" "SHLL_1D_TILES_PER_WARP=4,8"
} 2>&1 | tee $O/r2_07.log
