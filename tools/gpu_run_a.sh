#!/bin/bash
# GPU call A of round 2: parity of the rewritten kernels, sampler timings (both tuning variants), bench line, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -x -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/pytest_a.log
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "clustered or arena or closed_form or rebinding or shim or fuse or adamw or optimizer" -s 2>&1 | grep -E "clustered\[|arena |passed|failed|Error|error" | head -40 > gpurun_out/pytest_a_new.log
python tools/time_kernels.py --cases syaml16,sbase16,syaml64,hires1,sbase2 --dtypes bf16,f32 --json gpurun_out/time_v0.json > gpurun_out/time_v0.log 2>&1
TAMTR_MSDA_VARIANT=1 python tools/time_kernels.py --cases syaml16,syaml64 --dtypes bf16 --json gpurun_out/time_v1.json > gpurun_out/time_v1.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
tail -5 gpurun_out/bench_a.err > gpurun_out/bench_a.errtail
for k in fwd bwd; do
  ncu --set full --clock-control none --import-source on -k regex:msda_${k} --launch-skip 3 -c 1 -f -o gpurun_out/msda_r2_${k} \
      python tools/time_kernels.py --cases syaml16 --dtypes bf16 --iters 2 > gpurun_out/ncu_${k}.log 2>&1
done
ls -la gpurun_out
