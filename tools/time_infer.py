"""Inference latency / throughput of the MEH head (eval mode, bf16 autocast, CUDA events) at BASELINE.json config 5
(1280x1280 input: levels 320^2 / 160^2 / 80^2, 900 queries, 1 image per GPU-iteration) and at 640x640 batch 16, with the
VSSBlocks as identities and running for real."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tamtr_b200 import dp  # noqa: E402
from tamtr_b200.head import ManbaWorldDecoder  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def _flat(o):
    if isinstance(o, torch.Tensor):
        return [o]
    return [t for x in o for t in _flat(x)] if isinstance(o, (list, tuple)) else []


def main():
    torch.manual_seed(0)
    for name, B, sizes, nq in (("config 5: 1280^2, 900 queries, batch 1", 1, (320, 160, 80), 900),
                               ("config 5 at batch 8", 8, (320, 160, 80), 900),
                               ("640^2, 100 queries, batch 16", 16, (160, 80, 40), 100)):
        for vss in (False, True):
            m = ManbaWorldDecoder(10, [128, 256, 512], 512, nq, 4, 8, 3, vss=vss).cuda().eval()
            xs = [torch.randn(B, c, s, s, device="cuda").bfloat16() for c, s in zip((128, 256, 512), sizes)]
            text = torch.nn.functional.normalize(torch.randn(B, 10, 512, device="cuda"), dim=-1)

            def run():
                with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                    return m(xs, text)
            ms = timed(run, 5 if vss else 10)
            step = dp.HeadInferStep(m, (xs, text), autocast=torch.bfloat16)
            ref = run()
            got = step.run()
            same = all(torch.equal(a, b) for a, b in zip(_flat(ref), _flat(got)))
            ms_g = timed(step.run, 5 if vss else 20)
            print(f"{name:42s} vss={str(vss):5s}: eager {ms:7.2f} ms, graph {ms_g:7.2f} ms = {B / ms_g * 1e3:8.1f} img/s "
                  f"(same outputs: {same}; {step.launches_per_step} launches of ours)  "
                  f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
            del m, xs, step
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()


if __name__ == "__main__":
    main()
