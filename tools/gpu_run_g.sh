#!/bin/bash
# GPU call G of round 2: per-launch list of one VSSBlock step (bf16 autocast) at the three levels.
mkdir -p gpurun_out
python tools/profile_vss.py 128 160 > gpurun_out/profile_vss_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l0.csv python tools/profile_vss.py 128 160 > gpurun_out/ncu_g0.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l1.csv python tools/profile_vss.py 256 80 > gpurun_out/ncu_g1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l2.csv python tools/profile_vss.py 512 40 > gpurun_out/ncu_g2.log 2>&1
tail -2 gpurun_out/ncu_g2.log
