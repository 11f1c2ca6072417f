#!/bin/bash
O=gpurun_out; mkdir -p $O
export SWEEP_STEPS=512
{
echo "== STRICT o1 small grids: rows per chunk"
for n in 512 1024 2048; do echo "-- $n"; SWEEP_MODE=strict python tools/sweep2d.py o1:$n "SHLL_ROWS_PER_CHUNK=-,3,6,9,12,18,24"; done
echo "== STRICT o2 small grids"
for n in 256 512 1024; do echo "-- $n"; SWEEP_MODE=strict python tools/sweep2d.py o2:$n "SHLL_ROWS_PER_CHUNK=-,4,8,12,16,24,32"; done
echo "== FAST o1 small grids"
for n in 512 1024 2048; do echo "-- $n"; python tools/sweep2d.py o1:$n "SHLL_ROWS_PER_CHUNK=-,2,4,6,8,12,18"; done
echo "== FAST o2 small grids"
for n in 256 512 1024; do echo "-- $n"; python tools/sweep2d.py o2:$n "SHLL_ROWS_PER_CHUNK=-,4,8,12,16,24,32"; done
} 2>&1 | tee $O/r2_15.log
