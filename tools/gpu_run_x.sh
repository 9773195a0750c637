#!/bin/bash
# GPU call X of round 2 (final state): full GPU suite, full bench line, kernel breakdown of the graph-replayed step, launch list,
# smoke(), and three more quick bench lines kept whole (the step time is bimodal from process to process: 4.02 / 4.21 ms).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_x.log
tail -6 gpurun_out/pytest_x.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
tail -2 gpurun_out/bench_x.err; head -c 300 gpurun_out/bench_x.json; echo
timeout 300 python tools/profile_step.py --top 60 > gpurun_out/profile_step_x.log 2>&1
head -12 gpurun_out/profile_step_x.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_step_x.csv python bench.py --launch-list --steps 2 --warmup 1 > gpurun_out/ncu_x.log 2>&1
tail -1 gpurun_out/ncu_x.log
python tools/launch_summary.py gpurun_out/launches_step_x.csv > gpurun_out/launches_step_x_summary.txt; head -5 gpurun_out/launches_step_x_summary.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for i in 1 2 3; do timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_x.err | tail -1 > gpurun_out/bench_xq$i.json; python -c "import json; d=json.loads(open('gpurun_out/bench_xq$i.json').read()); print($i, round(d['ms_per_step'],4))"; done
