"""Device time of the background zero fill (csrc/core.cu: zero_fill_bg_kernel) for the S-yaml gradient arena (1.65 GB)
by CTA count, beside cudaMemsetAsync."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200 import _lib  # noqa: E402


def timed(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    n = 16 * 33600 * 1536 * 2
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = _lib.stream_ptr(buf.device)
    lib = _lib.lib()
    us = timed(lambda: lib.tamtr_memset_zero(buf.data_ptr(), n, st))
    print("cudaMemsetAsync        %8.1f us  %6.2f TB/s" % (us, n / us * 1e-6))
    for ctas in (8, 16, 24, 32, 48, 64, 96, 148, 296):
        us = timed(lambda: lib.tamtr_zero_fill_background(buf.data_ptr(), n, ctas, st))
        print("background, %3d CTAs   %8.1f us  %6.2f TB/s" % (ctas, us, n / us * 1e-6))


if __name__ == "__main__":
    main()
