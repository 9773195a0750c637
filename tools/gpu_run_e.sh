#!/bin/bash
# GPU call E of round 2: rewritten selective-scan kernels -- parity, then timing at the head's three levels.
mkdir -p gpurun_out
python -m pytest tests/test_vss_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_e.log
tail -8 gpurun_out/pytest_e.log
python tools/time_vss.py > gpurun_out/time_vss_e.log 2>&1
cat gpurun_out/time_vss_e.log | tail -8
