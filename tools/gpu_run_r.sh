#!/bin/bash
# GPU call R of round 2: full GPU suite with the folded encoder side, launch list of the head step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_r.log
tail -12 gpurun_out/pytest_r.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_r.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_r.log 2>&1
tail -1 gpurun_out/ncu_r.log
python tools/launch_summary.py gpurun_out/launches_step_r.csv | head -60
