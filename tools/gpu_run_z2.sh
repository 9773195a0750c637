#!/bin/bash
# GPU call Z2 of round 2: run-to-run spread of the two fill kernels with the fill forked behind the top-k, on ONE box; timeline.
mkdir -p gpurun_out
: > gpurun_out/ab_z2.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_z2.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1))" | tee -a gpurun_out/ab_z2.log; }
for i in 1 2 3; do
run TAMTR_ARENA_FILL_KERNEL=regs
run TAMTR_ARENA_FILL_KERNEL=bulk
done
run TAMTR_ARENA_FILL_KERNEL=regs
timeout 120 python tools/profile_timeline.py --device-targets > gpurun_out/timeline_z2.log 2>&1; head -30 gpurun_out/timeline_z2.log
tail -3 gpurun_out/bench_z2.err
