"""Device time of tamtr_rank_tokens (LayerNorm statistics + skinny score epilogue + max over classes) at the head shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import _lib  # noqa: E402

lib = _lib.lib()
B, Lv, d, nc = 16, 33600, 512, 10
E = torch.randn(B, Lv, d, device="cuda").bfloat16()
raw = torch.randn(B * Lv, 16, device="cuda")
enc_bias = torch.randn(d, device="cuda")
valid = (torch.rand(Lv, device="cuda") > 0.05).to(torch.uint8)
bw, sw, ck = (torch.randn(nc, device="cuda") for _ in range(3))
out = torch.empty(B * Lv, device="cuda")


def run():
    rc = lib.tamtr_rank_tokens(E.data_ptr(), raw.data_ptr(), enc_bias.data_ptr(), valid.data_ptr(), bw.data_ptr(), sw.data_ptr(),
                               ck.data_ptr(), out.data_ptr(), _lib.dtype_code(E), B, Lv, d, nc, 16, 1e-5, _lib.stream_ptr(E.device))
    assert rc == 0, lib.tamtr_last_error()


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"rank_tokens: {us:.1f} us, {(E.numel() * 2 + raw.numel() * 4) / us / 1e3:.0f} GB/s")
