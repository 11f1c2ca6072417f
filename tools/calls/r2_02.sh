#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== o1: chunk count vs waves (2368 resident warps; 69 tiles)"
python tools/sweep2d.py o1 "SHLL_NCHUNKS=103,137,171,172,205,206,228,240,241,274,275,343"
echo "== o1 without PDL at 228 / 240"
python tools/sweep2d.py o1 "SHLL_PDL=0" "SHLL_NCHUNKS=228,240"
echo "== o1 ring depth"
python tools/sweep2d.py o1 "SHLL_TMA_STAGES=2,3,4" "SHLL_NCHUNKS=206,240"
echo "== o2: register cap variants (acc_cfg) x chunk count (274 tiles)"
python tools/sweep2d.py o2 "SHLL_ACC_CFG=1,0,7,6,5" "SHLL_NCHUNKS=32"
python tools/sweep2d.py o2 "SHLL_ACC_CFG=1" "SHLL_NCHUNKS=19,26,32,38,39,45,52,64"
python tools/sweep2d.py o2 "SHLL_TMA_STAGES=2,3,4" "SHLL_NCHUNKS=32"
} 2>&1 | tee $O/r2_02_sweep.log
