#!/bin/bash
# GPU call M2 of round 2: does the step's mode (4.02 / 4.21 ms) change while one process keeps the GPU busy?  3 processes x 70 loops.
mkdir -p gpurun_out
for i in 1 2 3; do timeout 35 python tools/profile_timeline.py --device-targets --no-profile --loops 70 > gpurun_out/mode_m2_$i.log 2>&1; grep "^events" gpurun_out/mode_m2_$i.log | cut -c1-700; done
