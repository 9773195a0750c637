#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ab_w2.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_w.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1), d['clocks'])"; }
for i in 1 2 3 4; do run A=1 | tee -a gpurun_out/ab_w2.log; done
run TAMTR_LOWP_LAYER=0 | tee -a gpurun_out/ab_w2.log
run TAMTR_LOWP_LAYER=0 | tee -a gpurun_out/ab_w2.log
run TAMTR_ARENA_PREFILL=0 | tee -a gpurun_out/ab_w2.log
run TAMTR_ARENA_PREFILL=0 | tee -a gpurun_out/ab_w2.log
tail -3 gpurun_out/bench_w.err
