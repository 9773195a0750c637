#!/bin/bash
# GPU call V of round 2: decoder layer with bf16 operands from the add + LayerNorm kernels (TAMTR_LOWP_LAYER): tests, A/B.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layernorm_gpu.py tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_patch_gpu.py tests/test_fold_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_v.log
tail -6 gpurun_out/pytest_v.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_v.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1))"; }
run TAMTR_LOWP_LAYER=0 | tee gpurun_out/ab_v.log
run TAMTR_LOWP_LAYER=1 | tee -a gpurun_out/ab_v.log
run TAMTR_LOWP_LAYER=0 | tee -a gpurun_out/ab_v.log
run TAMTR_LOWP_LAYER=1 | tee -a gpurun_out/ab_v.log
tail -3 gpurun_out/bench_v.err
