#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== STRICT o1 4096^2: rows per chunk"
SWEEP_MODE=strict SWEEP_STEPS=100 python tools/sweep2d.py o1 "SHLL_ROWS_PER_CHUNK=6,9,12,18,24"
echo "== STRICT o2 2048x16384"
SWEEP_MODE=strict SWEEP_STEPS=50 python tools/sweep2d.py o2 "SHLL_ROWS_PER_CHUNK=12,16,24,32,64"
} 2>&1 | tee $O/r2_16.log
