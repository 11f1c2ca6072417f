#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== o1 occupancy variants (acc_cfg 1: 16 warps, 5: 18, 6: 20) x TMA store"
python tools/sweep2d.py o1 "SHLL_TMA_STORE=0" "SHLL_ACC_CFG=1,5,6" "SHLL_NCHUNKS=228,256"
python tools/sweep2d.py o1 "SHLL_TMA_STORE=1" "SHLL_ACC_CFG=1,5,6" "SHLL_NCHUNKS=228"
} 2>&1 | tee $O/r2_04.log
