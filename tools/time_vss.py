"""Device time of the selective scan and of whole VSSBlocks at the shapes of the MEH head (B=16, 640x640)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200 import _lib  # noqa: E402
from tamtr_b200.vss import VSSBlock, selective_scan  # noqa: E402


def timed(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = "cuda"
    for c, s in ((128, 160), (256, 80), (512, 40)):
        D, K, N, L = 2 * c, 4, 16, s * s
        u = torch.randn(B, K * D, L, device=dev, requires_grad=True)
        dt = (torch.randn(B, K * D, L, device=dev) - 2).requires_grad_()
        A = (-(0.5 + 15 * torch.rand(K * D, N, device=dev))).requires_grad_()
        Bm = torch.randn(B, K, N, L, device=dev, requires_grad=True)
        Cm = torch.randn(B, K, N, L, device=dev, requires_grad=True)
        Dv = torch.ones(K * D, device=dev, requires_grad=True)
        bias = torch.full((K * D,), -3.0, device=dev, requires_grad=True)
        with torch.no_grad():
            f_inf = timed(lambda: selective_scan(u, dt, A, Bm, Cm, Dv, bias))
        _lib.profile_enable(True)
        y = selective_scan(u, dt, A, Bm, Cm, Dv, bias)
        y.backward(torch.randn_like(y))
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile_enable(False)
        upd = B * K * D * L * N
        print(f"level c={c} {s}x{s}: scan fwd {f_inf:.2f} ms (inference), fwd+ckpt {prof['selective_scan_fwd'][0]:.2f} ms, "
              f"bwd {prof['selective_scan_bwd'][0]:.2f} ms; {upd / 1e9:.2f} G state updates -> "
              f"{upd / f_inf / 1e6:.0f} G updates/s fwd")
        del u, dt, Bm, Cm, y
        torch.cuda.empty_cache()
        blk = VSSBlock(hidden_dim=c, drop_path=0.0).to(dev)
        x = torch.randn(B, s, s, c, device=dev, requires_grad=True)

        def step():
            blk.zero_grad(set_to_none=True)
            x.grad = None
            blk(x).square().mean().backward()
        with torch.no_grad():
            t_f = timed(lambda: blk(x), 3)
        t_fb = timed(step, 3)
        print(f"           VSSBlock fwd {t_f:.2f} ms, fwd+bwd {t_fb:.2f} ms, peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
        del blk, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
