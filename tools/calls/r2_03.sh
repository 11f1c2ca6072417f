#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== parity of the FAST kernels with TMA stores"
timeout 900 python -m pytest tests/test_gpu_fast_parity.py tests/test_group.py -x -q 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fast or pdl or programmatic or golden" 2>&1 | tail -3
echo "== o1 TMA store on/off"
python tools/sweep2d.py o1 "SHLL_TMA_STORE=0,1" "SHLL_TMA_STAGES=2,3"
python tools/sweep2d.py o1 "SHLL_TMA_STORE=1" "SHLL_NCHUNKS=171,205,228,256,293"
echo "== o2 TMA store on/off"
python tools/sweep2d.py o2 "SHLL_TMA_STORE=0,1" "SHLL_TMA_STAGES=2,3"
} 2>&1 | tee $O/r2_03.log
