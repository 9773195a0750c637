"""Helper of tests/test_vss_gpu.py::test_scan_agrees_with_vllm_mamba_kernel, run in a subprocess (importing vllm is slow and
has side effects): our selective-scan forward against vLLM's `selective_scan_fn` -- a port of the mamba_ssm CUDA kernel,
the same kernel family as the `selective_scan_cuda_core` extension the reference calls (VManba/csms6s.py:257) but cannot
ship.  Library code, used here only as an independent implementation of the published recurrence.  Prints one JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import seeding
    from tamtr_b200.vss import selective_scan
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn
    except Exception as e:          # noqa: BLE001 -- any import problem means "not available here"
        print(json.dumps({"unavailable": f"{type(e).__name__}: {e}"[:300]}))
        return
    out = {}
    for name, (b, k, d, l) in {"small": (2, 4, 32, 70), "long": (1, 4, 64, 5000), "head_level2": (2, 4, 1024, 1600)}.items():
        n = 16
        u = seeding.seeded_tensor(41, "u", (b, k * d, l)).cuda()
        dt = (seeding.seeded_tensor(41, "dt", (b, k * d, l)) - 2.0).cuda()
        A = -(0.5 + 15.5 * seeding.seeded_uniform(41, "A", (k * d, n))).cuda()
        B = seeding.seeded_tensor(41, "B", (b, k, n, l)).cuda()
        C = seeding.seeded_tensor(41, "C", (b, k, n, l)).cuda()
        D = (1.0 + 0.2 * seeding.seeded_tensor(41, "D", (k * d,))).cuda()
        bias = (seeding.seeded_tensor(41, "bias", (k * d,)) - 3.0).cuda()
        with torch.no_grad():
            ours = selective_scan(u, dt, A, B, C, D, bias, True)
            state = torch.zeros(b, k * d, n, device="cuda")
            try:
                theirs = selective_scan_fn(u.clone(), state, dt.clone(), A, B, C, D, None, bias, True)
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"unavailable": f"{type(e).__name__}: {e}"[:300]}))
                return
        torch.cuda.synchronize()
        out[name] = ((ours.double() - theirs.double()).norm() / theirs.double().norm()).item()
    print(json.dumps({"rel_l2": out}))


if __name__ == "__main__":
    main()
