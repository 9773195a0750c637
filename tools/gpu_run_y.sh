#!/bin/bash
# GPU call Y of round 2: query-selection top-k kernel: its tests, the tests of everything that selects queries, timing, A/B.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_topk_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_y0.log
tail -4 gpurun_out/pytest_y0.log
timeout 120 python tools/time_topk.py 2>&1 | tee gpurun_out/time_topk.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_patch_gpu.py tests/test_fold_gpu.py tests/test_loss_gpu.py tests/test_cdn_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_y.log
tail -4 gpurun_out/pytest_y.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_y.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1))"; }
run TAMTR_TOPK=0 | tee gpurun_out/ab_y.log
run TAMTR_TOPK=1 | tee -a gpurun_out/ab_y.log
run TAMTR_TOPK=0 | tee -a gpurun_out/ab_y.log
run TAMTR_TOPK=1 | tee -a gpurun_out/ab_y.log
tail -3 gpurun_out/bench_y.err
