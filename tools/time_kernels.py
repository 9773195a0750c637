"""Kernel-only timings (CUDA events, L2 flushed between iterations) of the hot-path kernels.

    python tools/time_kernels.py [--iters 20] [--cases sbase16,syaml16,...] [--json out.json]
Algorithmic bytes follow SURVEY.md section 8(d) (compulsory value traffic + fp32 loc/attn + output)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamtr_b200  # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}

CASES = {
    # name: (B, Lq, H, Dh, base)
    "sbase2": (2, 300, 8, 32, 80),
    "sbase16": (16, 300, 8, 32, 80),
    "sbase16_q500": (16, 500, 8, 32, 80),
    "syaml16": (16, 300, 8, 64, 160),
    "syaml16_q100": (16, 100, 8, 64, 160),
    "syaml8": (8, 300, 8, 64, 160),
    "syaml64": (64, 300, 8, 64, 160),
    "hires1": (1, 900, 8, 64, 320),
}


def msda_bytes(B, Lq, H, Dh, shapes, sv, P=4):
    """Algorithmic bytes with the compulsory value traffic taken PER LEVEL: a level whose whole slab is smaller than what
    the gather would read from it (the 40x40 level at B=16, Lq=300) is counted once, the others by their gathered bytes."""
    d = H * Dh
    L = len(shapes)
    val = sum(min(B * h * w * d, B * Lq * H * P * 4 * Dh) for h, w in shapes)
    locw = B * Lq * H * L * P * 3 * 4
    fwd = val * sv + locw + B * Lq * d * sv
    bwd = B * Lq * d * sv + val * sv + 2 * locw + val * sv   # grad_value in value dtype (bf16 path: bf16 atomics)
    return fwd, bwd


def time_fn(fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for s, e in ev:
        flush.zero_()          # 512 MB write > 126 MB L2
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2] * 1e3, ts[0] * 1e3   # median, min in us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cases", default="sbase2,sbase16,syaml16,syaml16_q100,syaml64,hires1")
    ap.add_argument("--dtypes", default="bf16,f32")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name in args.cases.split(","):
        B, Lq, H, Dh, base = CASES[name]
        shapes = [[base, base], [base // 2, base // 2], [base // 4, base // 4]]
        Lv = sum(h * w for h, w in shapes)
        for dt in args.dtypes.split(","):
            dtype = torch.bfloat16 if dt == "bf16" else torch.float32
            g = torch.Generator(device=dev).manual_seed(0)
            value = torch.randn(B, Lv, H, Dh, device=dev, dtype=torch.float32, generator=g).to(dtype)
            cxcy = torch.rand(B, Lq, 1, 1, 1, 2, device=dev, generator=g)
            wh = 0.01 + 0.29 * torch.rand(B, Lq, 1, 1, 1, 2, device=dev, generator=g)
            loc = (cxcy + torch.randn(B, Lq, H, 3, 4, 2, device=dev, generator=g) * 2.0 / 4 * wh * 0.5).contiguous()
            attn = torch.softmax(torch.randn(B, Lq, H, 12, device=dev, generator=g), -1).view(B, Lq, H, 3, 4)
            gout = torch.randn(B, Lq, H * Dh, device=dev, generator=g).to(dtype)
            v = value.clone().requires_grad_()
            l = loc.clone().requires_grad_()
            a = attn.clone().requires_grad_()
            out = tamtr_b200.ms_deform_attn(v, shapes, l, a)
            fwd_us, fwd_min = time_fn(lambda: tamtr_b200.ms_deform_attn(value, shapes, loc, attn), args.iters, flush)

            def bwd():
                torch.autograd.grad(out, (v, l, a), gout, retain_graph=True)
            bwd_us, bwd_min = time_fn(bwd, args.iters, flush)
            fb, bb = msda_bytes(B, Lq, H, Dh, shapes, value.element_size())
            row = dict(case=name, dtype=dt, B=B, Lq=Lq, Dh=Dh, Lv=Lv, fwd_us=fwd_us, fwd_min_us=fwd_min,
                       bwd_us=bwd_us, bwd_min_us=bwd_min, fwd_MB=fb / 1e6, bwd_MB=bb / 1e6,
                       fwd_GBs=fb / fwd_us / 1e3, bwd_GBs=bb / bwd_us / 1e3,
                       fwd_frac=fb / fwd_us / 1e3 / PEAKS["hbm_gbs"], bwd_frac=bb / bwd_us / 1e3 / PEAKS["hbm_gbs"])
            rows.append(row)
            print(json.dumps(row), flush=True)
            del v, l, a, out, value, loc, attn, gout
            torch.cuda.empty_cache()
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
