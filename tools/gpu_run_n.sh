#!/bin/bash
# GPU call N of round 2: full bench line, ncu --set full of the scan kernels, launch lists (head step, level-0 VSSBlock step).
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err
tail -2 gpurun_out/bench_n.err; head -c 300 gpurun_out/bench_n.json; echo
python tools/profile_scan.py > gpurun_out/profile_scan_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sscan -s 2 -c 2 -f -o gpurun_out/sscan_r2 python tools/profile_scan.py > gpurun_out/ncu_n1.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_n.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_n2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l0_n.csv python tools/profile_vss.py 128 160 > gpurun_out/ncu_n3.log 2>&1
ls -la gpurun_out | tail -12
