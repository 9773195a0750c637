for d in 4 12; do echo "== project debug=$d"; TAMTR_TOK_DEBUG=$d python tools/time_tokgemm.py --iters 10 2>&1 | grep "tok_project" | cut -c1-150; done
for d in 0 32 34; do echo "== reduce debug=$d"; TAMTR_TOK_DEBUG=$d python tools/time_tokgemm.py --iters 10 --levels 0 2>&1 | grep "tok_reduce" | cut -c1-150; done
for i in 1 2 3; do timeout 200 python -m pytest tests/test_fold_gpu.py -q -k "weight_gradient or moments" 2>&1 | tail -1; done
