#!/bin/bash
# GPU call S of round 2: fold tests, quick bench, op-level and kernel-level profiles of the head step.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fold_gpu.py tests/test_step_gpu.py tests/test_patch_gpu.py -q -p no:cacheprovider 2>&1 | tail -6
timeout 300 python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_s_quick.json 2> gpurun_out/bench_s_quick.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_s_quick.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'], {k:round(v['avg_us'],1) for k,v in d.get('kernels',{}).items()})
PY
timeout 300 python tools/profile_ops.py --top 45 > gpurun_out/profile_ops_s.log 2>&1
timeout 300 python tools/profile_step.py --top 50 > gpurun_out/profile_step_s.log 2>&1
head -60 gpurun_out/profile_step_s.log
