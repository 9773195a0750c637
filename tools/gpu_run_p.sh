#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_selfattn_gpu.py tests/test_abi.py -q -x -p no:cacheprovider 2>&1 | tail -8
python tools/time_sdpa.py 2>&1 | grep -E "^ours|cudnn .*fwd\+bwd"
python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_patch_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -3
python bench.py --quick --steps 10 --warmup 3 > gpurun_out/bench_quick_p.json 2> gpurun_out/bench_quick_p.err
tail -1 gpurun_out/bench_quick_p.err; head -c 230 gpurun_out/bench_quick_p.json; echo
