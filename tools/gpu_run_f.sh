#!/bin/bash
# GPU call F of round 2: ncu --set full of the rewritten scan kernels at the head's largest level.
mkdir -p gpurun_out
python tools/profile_scan.py > gpurun_out/profile_scan_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sscan -s 2 -c 2 -f -o gpurun_out/sscan_r2a python tools/profile_scan.py > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
