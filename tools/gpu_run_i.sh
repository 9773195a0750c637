#!/bin/bash
# GPU call I of round 2: register-window depth-wise conv -- parity, per-level block timing, VSS-on bench line.
mkdir -p gpurun_out
python -m pytest tests/test_vss_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/pytest_i.log
tail -5 gpurun_out/pytest_i.log
python tools/profile_vss.py 128 160 > gpurun_out/profile_vss_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l0.csv python tools/profile_vss.py 128 160 > gpurun_out/ncu_i0.log 2>&1
python bench.py --vss --quick --steps 5 --warmup 3 > gpurun_out/bench_vss_i.json 2> gpurun_out/bench_vss_i.err
tail -2 gpurun_out/bench_vss_i.err; head -c 300 gpurun_out/bench_vss_i.json
