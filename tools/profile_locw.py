"""Kernel 3 (fused query projections + epilogue) at the bench shape, for ncu and CUDA-event timing."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200 import ops  # noqa: E402

M, C, H, L, P = 4800, 512, 8, 3, 4
torch.manual_seed(0)
q = torch.randn(1, M, C, device="cuda").bfloat16()
ref = torch.rand(1, M, 1, 4, device="cuda")
w_off = (torch.randn(H * L * P * 2, C, device="cuda") * 0.05).bfloat16()
w_att = (torch.randn(H * L * P, C, device="cuda") * 0.05).bfloat16()
b_off = torch.randn(H * L * P * 2, device="cuda")
b_att = torch.randn(H * L * P, device="cuda")
shapes = [[160, 160], [80, 80], [40, 40]]


def run():
    return ops.sampling_locations_and_weights(q, ref, w_off, b_off, w_att, b_att, shapes, H, L, P)


for fused in (True, False):
    ops.FUSED_PROJECTION = fused
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            run()
    g.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(10):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(f"fused={fused}: {a.elapsed_time(b) / 200 * 1e3:.2f} us per call (graph replay, includes the torch.cat / cast launches)")
