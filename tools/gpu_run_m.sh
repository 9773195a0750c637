#!/bin/bash
# GPU call M of round 2: timeline of the graph-replayed step in several fresh processes (the step time is bimodal per process).
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 100 python tools/profile_timeline.py --device-targets --gaps 6 > gpurun_out/timeline_m$i.log 2>&1; grep -m1 "^events" gpurun_out/timeline_m$i.log; grep -m2 "zero fill\|^step 1" gpurun_out/timeline_m$i.log; done
