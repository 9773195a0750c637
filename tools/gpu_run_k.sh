#!/bin/bash
# GPU call K of round 2: full GPU suite after the fused VSS tail / conv dgrad / gradient packing changes; VSS-on and
# headline quick bench lines; level-0 block launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=20 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_k.log
tail -6 gpurun_out/pytest_k.log
python bench.py --vss --quick --steps 5 --warmup 3 > gpurun_out/bench_vss_k.json 2> gpurun_out/bench_vss_k.err
tail -2 gpurun_out/bench_vss_k.err; head -c 260 gpurun_out/bench_vss_k.json; echo
python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_quick_k.json 2> gpurun_out/bench_quick_k.err
tail -2 gpurun_out/bench_quick_k.err; head -c 260 gpurun_out/bench_quick_k.json; echo
python tools/profile_vss.py 128 160 > gpurun_out/profile_vss_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vss_l0.csv python tools/profile_vss.py 128 160 > gpurun_out/ncu_k0.log 2>&1
