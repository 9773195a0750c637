"""One forward + backward selective scan at the head's largest level (B=16, K*D=1024, L=25600) for ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200.vss import selective_scan  # noqa: E402

B, K, D, N, L = 16, 4, 256, 16, 160 * 160
dev = "cuda"
u = torch.randn(B, K * D, L, device=dev, requires_grad=True)
dt = (torch.randn(B, K * D, L, device=dev) - 2).requires_grad_()
A = (-(0.5 + 15 * torch.rand(K * D, N, device=dev))).requires_grad_()
Bm = torch.randn(B, K, N, L, device=dev, requires_grad=True)
Cm = torch.randn(B, K, N, L, device=dev, requires_grad=True)
Dv = torch.ones(K * D, device=dev, requires_grad=True)
bias = torch.full((K * D,), -3.0, device=dev, requires_grad=True)
for _ in range(2):
    y = selective_scan(u, dt, A, Bm, Cm, Dv, bias)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")
