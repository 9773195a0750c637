"""Device time of tamtr_col_reduce2 (BatchNorm column statistics over the token tensor) at the head shapes."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import _lib
lib = _lib.lib()
B, Lv, d = 16, 33600, 512
a = torch.randn(B, Lv, d, device="cuda").bfloat16()
b = torch.randn(B, Lv, d, device="cuda").bfloat16()
for tok0, ntok in ((0, 25600), (25600, 6400), (32000, 1600)):
    ctas = lib.tamtr_col_reduce2_ctas(B, ntok)
    part = torch.empty(ctas, 2, d, device="cuda")
    def run():
        rc = lib.tamtr_col_reduce2(a.data_ptr(), b.data_ptr(), part.data_ptr(), _lib.dtype_code(a), B, Lv, d, tok0, ntok, _lib.stream_ptr(a.device))
        assert rc == 0, lib.tamtr_last_error()
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    gb = 2 * B * ntok * d * 2 / 1e9
    ref = (a[:, tok0:tok0+ntok].float().sum((0,1)), (a[:, tok0:tok0+ntok].float() * b[:, tok0:tok0+ntok].float()).sum((0,1)))
    got = part.double().sum(0)
    err = max(((got[0]-ref[0].double()).norm()/ref[0].double().norm()).item(), ((got[1]-ref[1].double()).norm()/ref[1].double().norm()).item())
    print(f"ntok={ntok}: {us:.1f} us, {gb/us*1e6:.0f} GB/s, ctas={ctas}, rel err {err:.2e}")
