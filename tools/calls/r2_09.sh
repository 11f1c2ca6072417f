#!/bin/bash
O=gpurun_out; mkdir -p $O
{
echo "== parity with the early-start protocol (1D tests, group tests)"
timeout 900 python -m pytest tests/test_group.py tests/test_gpu_parity.py tests/test_gpu_fast_parity.py -x -q -k "1d or persistent or graph or programmatic or group or golden" 2>&1 | tail -4
echo "== 1D FAST o2: early start on/off"
python tools/sweep1d.py 8388608 fast "SHLL_EARLY=0,1"
python tools/sweep1d.py 16777216 fast "SHLL_EARLY=0,1"
python tools/sweep1d.py 67108864 fast "SHLL_EARLY=0,1"
python tools/sweep1d.py 1048576 fast "SHLL_EARLY=0,1"
echo "== STRICT"
python tools/sweep1d.py 8388608 strict "SHLL_EARLY=0,1"
python tools/sweep1d.py 67108864 strict "SHLL_EARLY=0,1"
} 2>&1 | tee $O/r2_09.log
