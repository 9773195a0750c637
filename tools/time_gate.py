"""BASELINE.json config 3: max-sigmoid gate sweep (B=16, C=256, nh=8, hc=32, bf16), tcgen05 vs CUDA-core path."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import ops, _lib
dev = torch.device("cuda:0")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, iters=20):
    for _ in range(3): fn()
    evs = []
    for _ in range(iters):
        flush.zero_(); s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); evs.append((s, e))
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2] * 1e3
rows = []
cases = [(16, 8, hw, n) for hw in (400, 1600, 6400) for n in (10, 20, 40, 80)] + [(16, 2, 25600, 10), (16, 4, 6400, 10)]
for B, nh, HW, N in cases:
    side = int(HW ** 0.5)
    x = torch.randn(B, nh * 32, side, side, device=dev).bfloat16()
    g = torch.randn(B, N, nh, 32, device=dev) * 0.3
    bias = torch.zeros(nh, device=dev)
    t_tc = timeit(lambda: ops.max_sigmoid_gate(x, g, bias, nh, use_tensor_cores=True))
    t_cc = timeit(lambda: ops.max_sigmoid_gate(x, g, bias, nh, use_tensor_cores=False))
    byts = x.numel() * 2 + g.numel() * 4 + B * nh * HW * 5
    flops = 2.0 * B * HW * nh * 32 * N
    rows.append(dict(B=B, C=nh * 32, HW=HW, N=N, tc_us=t_tc, cuda_core_us=t_cc, MB=byts / 1e6, tc_GBs=byts / t_tc / 1e3,
                     tc_TFLOPs=flops / t_tc / 1e6, cc_GBs=byts / t_cc / 1e3))
    print(json.dumps(rows[-1]), flush=True)
json.dump(rows, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "gate_sweep_r1.json"), "w"), indent=1)
