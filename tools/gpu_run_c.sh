#!/bin/bash
# GPU call C of round 2: device-side CDN / loss kernels + fixed-shape step, full GPU suite, bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -3 gpurun_out/bench_c.err
head -c 600 gpurun_out/bench_c.json
