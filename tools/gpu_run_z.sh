#!/bin/bash
# GPU call Z of round 2: where the arena's zero fill is forked (right after the projection / behind the top-k) x which fill
# kernel (bulk stores from a shared tile / register stores without shared memory) x top-k kernel on / off, on ONE box.
mkdir -p gpurun_out
: > gpurun_out/ab_z.log
run() { env "$@" timeout 300 python bench.py --quick --steps 20 --warmup 5 2>> gpurun_out/bench_z.err | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],4), round(d['value'],1), round(d.get('e2e',{}).get('value',0),1))" | tee -a gpurun_out/ab_z.log; }
for i in 1 2; do
run TAMTR_TOPK=0 TAMTR_ARENA_DEFER=0
run TAMTR_TOPK=1 TAMTR_ARENA_DEFER=1
run TAMTR_TOPK=1 TAMTR_ARENA_DEFER=0 TAMTR_ARENA_FILL_CTAS=-148
run TAMTR_TOPK=1 TAMTR_ARENA_DEFER=1 TAMTR_ARENA_FILL_CTAS=-148
run TAMTR_TOPK=1 TAMTR_ARENA_DEFER=0 TAMTR_ARENA_FILL_CTAS=-296
done
tail -3 gpurun_out/bench_z.err
