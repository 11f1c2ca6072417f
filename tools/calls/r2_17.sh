#!/bin/bash
O=gpurun_out; mkdir -p $O
{
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "small_2d or ragged or golden or variant or geometry" 2>&1 | tail -3
export SWEEP_STEPS=512
for n in 512 1024 2048; do SWEEP_MODE=strict python tools/sweep2d.py o1:$n "SHLL_EARLY=1"; done
for n in 512 1024; do SWEEP_MODE=strict python tools/sweep2d.py o2:$n "SHLL_EARLY=1"; done
for n in 512 1024 2048; do python tools/sweep2d.py o1:$n "SHLL_EARLY=1"; done
} 2>&1 | tee $O/r2_17.log
