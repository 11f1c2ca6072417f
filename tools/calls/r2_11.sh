#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== parity (2D early start restricted)"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_parity.py tests/test_group.py -x -q 2>&1 | tail -3
echo "== 2D o2 FAST 2048x16384: 0 off, 1 auto, 2 forced"
python tools/sweep2d.py o2 "SHLL_EARLY=0,1"
echo "== 2D o1 FAST 4096^2 (auto = off)"
python tools/sweep2d.py o1 "SHLL_EARLY=0,1,2"
echo "== STRICT"
SWEEP_MODE=strict SWEEP_STEPS=100 python tools/sweep2d.py o1 "SHLL_EARLY=0,1"
SWEEP_MODE=strict SWEEP_STEPS=60 python tools/sweep2d.py o2 "SHLL_EARLY=0,1"
echo "== small grids stay as they were"
for n in 256 1024; do SHLL_GRAPH=0 python tools/sweep2d.py o1:$n "SHLL_EARLY=0,1"; done
} 2>&1 | tee $O/r2_11.log
