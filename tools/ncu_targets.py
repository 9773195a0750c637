"""Small driver for ncu captures: runs rank_tokens, the tcgen05 gate and the token-major BN kernels at bench shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
B, Lv, d, nc = 16, 33600, 512, 10
feats = torch.randn(B, Lv, d, device=dev).bfloat16()
lin = torch.nn.Linear(d, d).to(dev); ln = torch.nn.LayerNorm(d).to(dev); sc = torch.nn.Linear(d, nc).to(dev)
valid = torch.ones(Lv, dtype=torch.uint8, device=dev)
for _ in range(3):
    r = ops.rank_tokens(feats, valid, lin, ln, sc)
x = torch.randn(16, 256, 80, 80, device=dev).bfloat16(); g = torch.randn(16, 80, 8, 32, device=dev) * 0.3; bias = torch.zeros(8, device=dev)
for n in (10, 80):
    for _ in range(2):
        ops.max_sigmoid_gate(x, g[:, :n].contiguous(), bias, 8, use_tensor_cores=True)
torch.cuda.synchronize()
print("ok", r.shape)
