#!/bin/bash
mkdir -p gpurun_out
python tools/profile_ops.py --top 70 > gpurun_out/profile_ops.log 2>&1
tail -130 gpurun_out/profile_ops.log
