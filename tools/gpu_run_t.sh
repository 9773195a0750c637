#!/bin/bash
# GPU call T of round 2: full GPU suite, full bench line, kernel breakdown of the graph-replayed step, launch list, and one
# ncu --set full capture of the folded-projection kernels at level 0.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_t.log
tail -6 gpurun_out/pytest_t.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
tail -2 gpurun_out/bench_t.err; head -c 300 gpurun_out/bench_t.json; echo
timeout 300 python tools/profile_step.py --top 60 > gpurun_out/profile_step_t.log 2>&1
head -30 gpurun_out/profile_step_t.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_t.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_t.log 2>&1
tail -1 gpurun_out/ncu_t.log
python tools/launch_summary.py gpurun_out/launches_step_t.csv > gpurun_out/launches_step_t_summary.txt; head -5 gpurun_out/launches_step_t_summary.txt
ncu --set full --clock-control none --import-source on -k regex:tok_ -c 16 -o gpurun_out/tokgemm_t python tools/time_tokgemm.py --iters 1 --levels 0 > gpurun_out/ncu_tok_t.log 2>&1
tail -2 gpurun_out/ncu_tok_t.log
