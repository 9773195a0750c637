#!/bin/bash
# GPU call U of round 2: the gradient arena's zero fill forked to a side stream (TAMTR_ARENA_PREFILL) as a full-grid memset
# node or as a few-CTA bulk-store kernel (TAMTR_ARENA_FILL_CTAS): tests it touches, the fill kernel alone, A/B of the step.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fold_gpu.py tests/test_step_gpu.py tests/test_abi.py -q -x -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest_u.log
tail -4 gpurun_out/pytest_u.log
timeout 120 python tools/time_zero_fill.py 2>&1 | tee gpurun_out/time_zero_fill.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_u.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1))"; }
run TAMTR_ARENA_PREFILL=0 | tee gpurun_out/ab_u.log
for c in 0 16 32 64 148; do run TAMTR_ARENA_PREFILL=1 TAMTR_ARENA_FILL_CTAS=$c | tee -a gpurun_out/ab_u.log; done
run TAMTR_ARENA_PREFILL=0 | tee -a gpurun_out/ab_u.log
tail -3 gpurun_out/bench_u.err
