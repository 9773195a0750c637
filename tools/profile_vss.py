"""One VSSBlock forward + backward at one level of the head (default: the largest, B=16, 160x160, C=128) -- for a per-launch ncu list."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200.vss import VSSBlock  # noqa: E402

C, S = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 160)      # level: channels, map side
torch.manual_seed(0)
blk = VSSBlock(hidden_dim=C, drop_path=0.0).cuda()
x = torch.randn(16, S, S, C, device="cuda", requires_grad=True)
for _ in range(2):
    blk.zero_grad(set_to_none=True)
    x.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = blk(x)
    y.float().square().mean().backward()
torch.cuda.synchronize()
print("ok")
