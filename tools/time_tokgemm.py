"""Kernel-only timings of the folded-projection kernels (csrc/tokgemm.cu) at TAM-TR shapes, per pyramid level.

    python tools/time_tokgemm.py [--iters 20] [--batch 16] [--json out.json]
CUDA events, a 512 MB write between iterations (> 126 MB L2).  Algorithmic bytes: project = X + W_fold + every output
column once; reduce = A + X once (+ the fp32 partials)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamtr_b200  # noqa: E402
from tamtr_b200 import fold  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def time_fn(fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--json", default=None)
    ap.add_argument("--levels", default="0,1,2")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, d, n_layers = args.batch, 512, 3
    N0, N1, NT = n_layers * d, d, 16
    chans, sizes = (128, 256, 512), (160, 80, 40)
    Lv = sum(s * s for s in sizes)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out0 = torch.empty(B, Lv, N0, dtype=torch.bfloat16, device=dev)
    out1 = torch.empty(B, Lv, N1, dtype=torch.bfloat16, device=dev)
    raw = torch.empty(B, Lv, NT, dtype=torch.float32, device=dev)
    scores = torch.empty(B, Lv, dtype=torch.float32, device=dev)
    grad = torch.randn(B, Lv, N0, device=dev).bfloat16()
    rows = []
    start = 0
    for l, (C, s) in enumerate(zip(chans, sizes)):
        HW = s * s
        if str(l) in args.levels.split(","):
            x = torch.randn(B, C, s, s, device=dev).bfloat16()
            w = (torch.randn(N0 + N1 + NT, C, device=dev) / C ** 0.5).bfloat16()
            bias = torch.randn(N0 + N1 + NT, device=dev)
            t = time_fn(lambda: fold._kernel_project(x, w, bias, out0, out1, raw, start, N0, N1, NT), args.iters, flush)
            by = x.numel() * 2 + w.numel() * 2 + B * HW * ((N0 + N1) * 2 + NT * 4)
            rows.append({"kernel": "tok_project", "level": l, "C": C, "tokens": B * HW, "us": t, "bytes": by,
                         "gbs": by / t / 1e3, "frac": by / t / 1e3 / PEAK, "tflops": 2.0 * B * HW * C * (N0 + N1 + NT) / t / 1e6})
            consts = torch.randn(2 + 3 * NT, device=dev)
            consts[1] = consts[1].abs() + 600.0
            valid = torch.ones(Lv, dtype=torch.uint8, device=dev)
            t = time_fn(lambda: fold._kernel_project(x, w, bias, out0, None, None, start, N0, N1, NT,
                                                     (scores, valid, consts, 10, 1e-5)), args.iters, flush)
            by = x.numel() * 2 + w.numel() * 2 + B * HW * (N0 * 2 + 4)
            rows.append({"kernel": "tok_project_rank", "level": l, "C": C, "tokens": B * HW, "us": t, "bytes": by,
                         "gbs": by / t / 1e3, "frac": by / t / 1e3 / PEAK, "tflops": 2.0 * B * HW * C * (N0 + N1 + NT) / t / 1e6})
            t = time_fn(lambda: fold._kernel_reduce(x, HW, C * HW, False, x, C), args.iters, flush)
            by = x.numel() * 2
            rows.append({"kernel": "tok_reduce(moments)", "level": l, "C": C, "tokens": B * HW, "us": t, "bytes": by,
                         "gbs": by / t / 1e3, "frac": by / t / 1e3 / PEAK})
            a = grad[:, start:]
            t = time_fn(lambda: fold._kernel_reduce(a, N0, Lv * N0, True, x, N0), args.iters, flush)
            by = x.numel() * 2 + B * HW * N0 * 2
            rows.append({"kernel": "tok_reduce(wgrad)", "level": l, "C": C, "tokens": B * HW, "us": t, "bytes": by,
                         "gbs": by / t / 1e3, "frac": by / t / 1e3 / PEAK, "tflops": 2.0 * B * HW * C * N0 / t / 1e6})
        start += HW
    for r in rows:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})
    if args.json:
        json.dump({"peak_gbs": PEAK, "batch": B, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
