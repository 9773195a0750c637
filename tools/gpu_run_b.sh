#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_b.log
python tools/exp_layout.py > gpurun_out/exp_layout.json 2> gpurun_out/exp_layout.err
tail -3 gpurun_out/exp_layout.err
cat gpurun_out/exp_layout.json
tail -5 gpurun_out/pytest_b.log
