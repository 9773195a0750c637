"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference in the build container.

    python -m oracle.make_goldens [names...]

/root/reference is pure PyTorch, so "outputs of the reference itself run here" are produced by importing it
(oracle/reference_loader.py) on CPU, fp32.  Inputs are regenerated from seeds by the tests (CPU generators are
machine-independent), so the fixtures only hold seeds, shapes and reference OUTPUTS (full tensors when small,
otherwise a fixed random subset + norms).  Every fixture records the reference file:line that produced it.
"""
import os
import sys

import numpy as np
import torch

from . import msda, reference_loader, seeding

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _subset(t, n, seed):
    """Fixed random subset of a big tensor: (flat indices int64, values)."""
    flat = t.detach().reshape(-1)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(flat.numel(), generator=g)[:n].sort().values
    return idx, flat[idx].clone()


def _save(name, obj):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e3:.1f} kB)")


# ------------------------------------------------------------------------------------------------ core op
CORE_CASES = {
    # name: (seed, B, Lq, H, Dh, shapes, oob_frac)
    "tiny_nonsquare": (11, 2, 17, 4, 8, [[9, 13], [5, 7], [3, 4]], 0.30),
    "small_dh32": (12, 2, 50, 8, 32, [[20, 20], [10, 10], [5, 5]], 0.25),
    "small_dh64": (13, 2, 40, 8, 64, [[24, 16], [12, 8], [6, 4]], 0.25),
    "small_dh16_L4": (14, 1, 33, 4, 16, [[16, 16], [8, 8], [4, 4], [2, 2]], 0.25),
    "sbase_b2": (15, 2, 300, 8, 32, [[80, 80], [40, 40], [20, 20]], 0.25),       # BASELINE.json config 1 shapes
    "syaml_b1": (16, 1, 300, 8, 64, [[160, 160], [80, 80], [40, 40]], 0.25),     # TAMTR.yaml:67 shapes
}


def gen_core(ns):
    ref = ns.multi_scale_deformable_attn_pytorch   # ultralytics/nn/modules/utils.py:42
    out = {"source": "ultralytics/nn/modules/utils.py:42-89 multi_scale_deformable_attn_pytorch, torch "
                     + torch.__version__ + " CPU fp32", "cases": {}}
    for name, (seed, B, Lq, H, Dh, shapes, oob) in CORE_CASES.items():
        value, loc, attn, grad_out = msda.make_inputs(seed, B, Lq, H, Dh, shapes, oob_frac=oob)
        value.requires_grad_(), loc.requires_grad_(), attn.requires_grad_()
        o = ref(value, shapes, loc, attn)
        o.backward(grad_out)
        case = dict(seed=seed, B=B, Lq=Lq, H=H, Dh=Dh, shapes=shapes, oob_frac=oob, P=4,
                    grad_loc=loc.grad.clone(), grad_attn=attn.grad.clone(),
                    out_norm=o.detach().double().norm().item(),
                    grad_value_norm=value.grad.double().norm().item(),
                    grad_value_sum=value.grad.double().sum().item())
        if value.numel() <= 300_000:
            case["out"] = o.detach().clone()
            case["grad_value"] = value.grad.clone()
        else:
            case["out"] = o.detach().clone()
            case["grad_value_subset"] = _subset(value.grad, 20000, seed + 1000)
        out["cases"][name] = case
    _save("msda_core", out)


def gen_index_probe(ns):
    """Bit-exact index contract: identity images expose the bilinear weights, hence the bits
    of ix, hence floor(ix) and the in-bounds flags, of torch's grid_sampler as the reference calls it."""
    ref = ns.multi_scale_deformable_attn_pytorch
    out = {"source": "utils.py:74-78 F.grid_sample(bilinear, zeros, align_corners=False) through "
                     "multi_scale_deformable_attn_pytorch on identity images; torch " + torch.__version__ + " CPU",
           "cases": {}}
    for W in (20, 40, 80, 160, 320, 13, 7):
        k = np.arange(-1, W + 2, dtype=np.float64)
        pts = []
        for base in ((k + 0.5) / W, k / W, (k + 0.25) / W):
            b32 = base.astype(np.float32)
            for d in (-2, -1, 0, 1, 2):
                v = b32.copy()
                for _ in range(abs(d)):
                    v = np.nextafter(v, np.float32(np.inf if d > 0 else -np.inf), dtype=np.float32)
                pts.append(v)
        rnd = (np.random.RandomState(W).rand(1500).astype(np.float32) * np.float32(1.2) - np.float32(0.1))
        pts = torch.from_numpy(np.concatenate(pts + [rnd]))
        n = pts.numel()
        # x-direction probe: level [1, W], y fixed at 0.5 (iy == 0 exactly -> weight 1 on row 0)
        loc = torch.zeros(1, n, 1, 1, 1, 2)
        loc[0, :, 0, 0, 0, 0] = pts
        loc[0, :, 0, 0, 0, 1] = 0.5
        ox = ref(torch.eye(W).view(1, W, 1, W), [[1, W]], loc, torch.ones(1, n, 1, 1, 1))[0]     # [n, W]
        # y-direction probe: level [W, 1]
        loc_y = torch.zeros(1, n, 1, 1, 1, 2)
        loc_y[0, :, 0, 0, 0, 1] = pts
        loc_y[0, :, 0, 0, 0, 0] = 0.5
        oy = ref(torch.eye(W).view(1, W, 1, W), [[W, 1]], loc_y, torch.ones(1, n, 1, 1, 1))[0]
        out["cases"][W] = dict(pts=pts, x_sparse=ox.to_sparse_coo(), y_sparse=oy.to_sparse_coo())
    _save("msda_index_probe", out)


GENERATORS = {"core": gen_core, "index_probe": gen_index_probe}


def main(argv):
    ns = reference_loader.hot_path()
    torch.manual_seed(0)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    names = argv or list(GENERATORS)
    for n in names:
        GENERATORS[n](ns)


if __name__ == "__main__":
    main(sys.argv[1:])
