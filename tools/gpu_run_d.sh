#!/bin/bash
# GPU call D of round 2: state check after container re-creation -- full GPU suite, bench line, launch list of the step.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_d.log
tail -5 gpurun_out/pytest_d.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err
tail -3 gpurun_out/bench_d.err
head -c 400 gpurun_out/bench_d.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_d.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_d.log 2>&1
python tools/time_vss.py > gpurun_out/time_vss_d.log 2>&1
tail -20 gpurun_out/time_vss_d.log
