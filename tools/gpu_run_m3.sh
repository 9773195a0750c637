#!/bin/bash
# GPU call M3 of round 2: clocks / performance state sampled every 100 ms while one process runs 110 loops of 20 replays
# (the step flips from 4.20 to 4.02 ms some seconds into a busy period: which clock moves with it?).
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,clocks.gr,clocks.video,pstate,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader -lms 100 > gpurun_out/mode_m3_smi.csv 2>&1 &
SMI=$!
timeout 30 python tools/profile_timeline.py --device-targets --no-profile --loops 110 > gpurun_out/mode_m3.log 2>&1
kill $SMI
grep -c . gpurun_out/mode_m3_smi.csv; tail -3 gpurun_out/mode_m3.log | cut -c1-200
