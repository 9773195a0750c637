"""Shared helpers for the parity tests."""
import os

import torch

from oracle import seeding

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_l2(x, y):
    x, y = x.detach().double().cpu(), y.detach().double().cpu()
    return ((x - y).norm() / y.norm().clamp_min(1e-30)).item()


def subset_err(t, subset):
    """relative L2 error of tensor `t` on a golden (indices, values) subset."""
    idx, vals = subset
    return rel_l2(t.detach().reshape(-1).cpu()[idx.long()], vals)


def check_full_or_subset(t, case, key, tol):
    """Golden stores `key` (full) for small cases or `key_subset` + `key_norm` for full-size ones."""
    if key in case and case[key] is not None:
        assert rel_l2(t, case[key]) < tol, (key, rel_l2(t, case[key]))
    else:
        assert subset_err(t, case[key + "_subset"]) < tol, (key, subset_err(t, case[key + "_subset"]))
        n = t.detach().double().norm().item()
        assert abs(n - case[key + "_norm"]) <= tol * case[key + "_norm"] + 1e-12, key


def probe_loss(out, seed, name):
    return (out * seeding.seeded_tensor(seed, name, out.shape).to(out.device, out.dtype)).sum()


def filled_state_dict(module, seed, manifest=None):
    """Apply the deterministic fill to `module` and (optionally) check its keys/shapes against the reference's."""
    got = seeding.seeded_fill(module, seed)
    if manifest is not None:
        ref = {k: tuple(v) for k, v in manifest.items() if not k.startswith("VSSBlocks.")}
        assert got == ref, ("state_dict keys/shapes differ from the reference",
                            set(got.items()) ^ set(ref.items()))
    return {k: v.detach().clone() for k, v in module.state_dict().items()}


def align_queries(ours_boxes, ours_scores, gold_boxes, gold_scores):
    """Query order comes from a top-k over thousands of scores: two implementations that agree to 1e-6 can still swap
    near-tied neighbours.  Match rows through their encoder outputs (boxes alone are not distinct: many saturate at 1.0
    with the seeded weights).  Returns (index of the golden row for each of our rows, matched mask)."""
    a = torch.cat([ours_boxes, ours_scores], -1).detach().double().cpu()
    b = torch.cat([gold_boxes, gold_scores], -1).detach().double().cpu()
    dist, idx = torch.cdist(a, b).min(-1)                                          # [B, nq]
    return idx, dist < 1e-3


def gather_rows(t, idx):
    return torch.gather(t, 1, idx.unsqueeze(-1).expand(-1, -1, t.shape[-1]))
