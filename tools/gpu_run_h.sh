#!/bin/bash
# GPU call H of round 2: fused SS2D function -- parity (VSS tests), then per-launch list of one VSSBlock step at level 0.
mkdir -p gpurun_out
python -m pytest tests/test_vss_gpu.py tests/test_patch_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_h.log
tail -12 gpurun_out/pytest_h.log
python tools/profile_vss.py 128 160 > gpurun_out/profile_vss_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l0.csv python tools/profile_vss.py 128 160 > gpurun_out/ncu_h0.log 2>&1
tail -2 gpurun_out/profile_vss_plain.log
python tools/time_vss.py > gpurun_out/time_vss_h.log 2>&1
grep VSSBlock gpurun_out/time_vss_h.log
