#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== parity with the 2D early-start protocol"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_parity.py tests/test_group.py tests/test_host_programs.py -x -q 2>&1 | tail -4
echo "== 2D o1 4096^2: early start on/off"
python tools/sweep2d.py o1 "SHLL_EARLY=0,1"
python tools/sweep2d.py o1 "SHLL_EARLY=1" "SHLL_EARLY_BLOCKS=1184,2368,4736,9999999"
echo "== 2D o2 2048x16384"
python tools/sweep2d.py o2 "SHLL_EARLY=0,1"
echo "== reference sizes (graph off so that launches are programmatic)"
for n in 256 512 1024 2048; do SHLL_GRAPH=0 python tools/sweep2d.py o1:$n "SHLL_EARLY=0,1"; done
for n in 256 512 1024; do SHLL_GRAPH=0 python tools/sweep2d.py o2:$n "SHLL_EARLY=0,1"; done
} 2>&1 | tee $O/r2_10.log
