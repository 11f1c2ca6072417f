#!/bin/bash
O=gpurun_out; mkdir -p $O
{
python tools/sweep2d.py o1 "SHLL_FUSE2=1" "SHLL_ROWS_PER_CHUNK=28,32,40,44,48,52,56,64,72,84,100"
echo "== small grids, fused: rows per chunk"
export SWEEP_STEPS=512
for n in 512 1024 2048; do echo "-- $n"; python tools/sweep2d.py o1:$n "SHLL_ROWS_PER_CHUNK=-,4,8,12,16,20,28"; done
echo "-- 1024 unfused for reference"; python tools/sweep2d.py o1:1024 "SHLL_FUSE2=0"
} 2>&1 | tee $O/r2_21.log
