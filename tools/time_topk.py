"""tamtr_topk_rows against torch.topk on the query-selection shapes, CUDA-graph replays timed with events.
    python tools/time_topk.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import ops

def timed(fn, iters=50):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters

for name, B, n, k, scale, shift in [("640^2 train, spread scores", 16, 33600, 300, 3.0, -1.0),
                                    ("640^2 train, scores around the -4.6 bias", 16, 33600, 300, 0.3, -4.6),
                                    ("config 1", 2, 8400, 300, 1.0, 0.0), ("1280^2 inference", 1, 134400, 900, 0.3, -4.6)]:
    s = (torch.randn(B, n, generator=torch.Generator().manual_seed(1)) * scale + shift).cuda()
    ours = timed(lambda: ops.topk_rows(s, k))
    lib = timed(lambda: torch.topk(s, k, dim=1).indices)
    print(f"{name:45s} B={B:3d} n={n:6d} k={k:4d}  tamtr_topk_rows {ours:7.1f} us   torch.topk {lib:7.1f} us")
