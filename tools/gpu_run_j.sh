#!/bin/bash
mkdir -p gpurun_out
python tools/profile_ops.py --top 5 > gpurun_out/profile_ops.log 2>&1
tail -115 gpurun_out/profile_ops.log
