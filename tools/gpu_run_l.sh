#!/bin/bash
# GPU call L of round 2: pyramid levels' VSSBlocks on parallel streams (graph branches) vs one stream.
mkdir -p gpurun_out
python -m pytest tests/test_modules_gpu.py tests/test_step_gpu.py tests/test_patch_gpu.py -q -x -p no:cacheprovider 2>&1 | tail -5
for p in 1 0; do
  TAMTR_VSS_PARALLEL=$p python bench.py --vss --quick --steps 5 --warmup 3 > gpurun_out/bench_vss_par$p.json 2> gpurun_out/bench_vss_par$p.err
  tail -1 gpurun_out/bench_vss_par$p.err; head -c 230 gpurun_out/bench_vss_par$p.json; echo
done
