#!/bin/bash
# GPU call Q of round 2: full GPU suite, full bench line, launch list of the head step (summary only is committed).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -5 > gpurun_out/pytest_q.log
tail -3 gpurun_out/pytest_q.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
tail -2 gpurun_out/bench_q.err; head -c 260 gpurun_out/bench_q.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_q.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_q.log 2>&1
tail -1 gpurun_out/ncu_q.log
