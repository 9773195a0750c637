"""CPU tests: pin the oracle against golden vectors produced by the reference itself (oracle/make_goldens.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import msda
from helpers import check_full_or_subset


def rel_l2(x, y):
    return ((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def core(golden_dir):
    return torch.load(os.path.join(golden_dir, "msda_core.pt"), weights_only=False)


@pytest.fixture(scope="module")
def probe(golden_dir):
    return torch.load(os.path.join(golden_dir, "msda_index_probe.pt"), weights_only=False)


CASES = ["tiny_nonsquare", "small_dh32", "small_dh64", "small_dh16_L4", "sbase_b2", "syaml_b1"]


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_golden(core, name):
    c = core["cases"][name]
    value, loc, attn, grad_out = msda.make_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"],
                                                  oob_frac=c["oob_frac"])
    out = msda.forward_c(value, c["shapes"], loc, attn)
    gv, gl, ga = msda.backward_c(grad_out, value, c["shapes"], loc, attn)
    for t, key in ((out, "out"), (gv, "grad_value"), (gl, "grad_loc"), (ga, "grad_attn")):
        check_full_or_subset(t, c, key, 5e-6)                # fp32 tolerance of the north star is 1e-4


@pytest.mark.parametrize("name", ["tiny_nonsquare", "small_dh16_L4"])
def test_explicit_and_gridsample_restatements(core, name):
    c = core["cases"][name]
    value, loc, attn, grad_out = msda.make_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"],
                                                  oob_frac=c["oob_frac"])
    for fn in (msda.msda_explicit_torch, msda.msda_gridsample_torch):
        v, l, a = (t.clone().requires_grad_() for t in (value, loc, attn))
        out = fn(v, c["shapes"], l, a)
        out.backward(grad_out)
        assert rel_l2(out.detach(), c["out"]) < 2e-6
        assert rel_l2(v.grad, c["grad_value"]) < 2e-6
        assert rel_l2(l.grad, c["grad_loc"]) < 5e-6
        assert rel_l2(a.grad, c["grad_attn"]) < 2e-6


def test_c_backward_matches_fp64_autograd():
    """The hand-written backward in msda_core.c against autograd of the explicit fp64 restatement."""
    shapes = [[7, 5], [4, 3]]
    value, loc, attn, grad_out = msda.make_inputs(3, 2, 9, 2, 8, shapes, P=3, oob_frac=0.3)
    v, l, a = (t.double().requires_grad_() for t in (value, loc, attn))
    out = msda.msda_explicit_torch(v, shapes, l, a)
    out.backward(grad_out.double())
    gv, gl, ga = msda.backward_c(grad_out, value, shapes, loc, attn)
    assert rel_l2(gv, v.grad) < 1e-6 and rel_l2(gl, l.grad) < 1e-5 and rel_l2(ga, a.grad) < 1e-6


@pytest.mark.parametrize("W", [20, 40, 80, 160, 320, 13, 7])
def test_index_math_bit_exact_vs_reference(probe, W):
    """Identity images make the output equal to the bilinear weights, i.e. expose every bit of ix / iy and the
    in-bounds decisions.  EXACT equality with what the reference (torch grid_sample) produced."""
    c = probe["cases"][W]
    pts = c["pts"]
    n = pts.numel()
    eye = torch.eye(W).view(1, W, 1, W)
    ones = torch.ones(1, n, 1, 1, 1)
    loc = torch.zeros(1, n, 1, 1, 1, 2)
    loc[0, :, 0, 0, 0, 0] = pts
    loc[0, :, 0, 0, 0, 1] = 0.5
    ox = msda.forward_c(eye, [[1, W]], loc, ones)[0]
    assert torch.equal(ox, c["x_sparse"].to_dense())
    loc_y = torch.zeros(1, n, 1, 1, 1, 2)
    loc_y[0, :, 0, 0, 0, 1] = pts
    loc_y[0, :, 0, 0, 0, 0] = 0.5
    oy = msda.forward_c(eye, [[W, 1]], loc_y, ones)[0]
    assert torch.equal(oy, c["y_sparse"].to_dense())
    # corners export agrees with the weights' support
    x0, _, inb = msda.corners_c(loc, [[1, W]])
    dense = c["x_sparse"].to_dense()
    x0 = x0.view(-1).long()
    for k, col in ((0, x0), (1, x0 + 1)):
        valid = inb.view(n, 4)[:, k].bool()
        assert ((col >= 0) & (col < W))[valid].all()
        # every non-zero weight of the reference sits on an in-bounds corner of the oracle
    nz = dense.nonzero()
    rows, cols = nz[:, 0], nz[:, 1]
    assert (((cols == x0[rows]) & inb.view(n, 4)[rows, 0].bool()) |
            ((cols == x0[rows] + 1) & inb.view(n, 4)[rows, 1].bool())).all()


def test_init_state_known_answer():
    """KAT from transformer.py:234-250: all 12 attention weights equal 1/12 -> output is the mean of 12 taps."""
    shapes = [[6, 6], [3, 3]]
    value, loc, _, _ = msda.make_inputs(5, 1, 4, 2, 8, shapes, P=2)
    attn = torch.full((1, 4, 2, 2, 2), 0.25)
    out = msda.forward_c(value, shapes, loc, attn)
    per_tap = []
    for l in range(2):
        for p in range(2):
            onehot = torch.zeros_like(attn)
            onehot[:, :, :, l, p] = 1.0
            per_tap.append(msda.forward_c(value, shapes, loc, onehot))
    assert torch.allclose(out, torch.stack(per_tap).mean(0), atol=1e-6)


def test_out_of_range_locations_give_zero():
    shapes = [[4, 4]]
    value = torch.randn(1, 16, 1, 8)
    loc = torch.tensor([[-0.5, 0.5], [1.5, 0.5], [0.5, -0.3], [0.5, 1.3], [float("nan"), 0.5], [1e30, 0.5]])
    loc = loc.view(1, 6, 1, 1, 1, 2)
    out = msda.forward_c(value, shapes, loc, torch.ones(1, 6, 1, 1, 1))
    assert torch.equal(out, torch.zeros_like(out))


def test_topk_oracle_pinned_to_the_reference_call():
    """oracle/topk.py against the reference's own call, torch.topk(scores, nq, dim=1).indices (head.py:1240, :437): equal
    indices wherever the scores are distinct, equal values always; its documented tie order (lower index first), NaN first,
    -0.0 == +0.0."""
    from oracle import topk
    g = torch.Generator().manual_seed(3)
    for B, n, k in [(2, 8400, 300), (3, 33600, 300), (1, 1000, 1000), (4, 17, 1)]:
        scores = torch.stack([torch.randperm(n, generator=g).float() for _ in range(B)]) * (8.0 / n) - 4.6   # distinct
        assert all(len(set(r.tolist())) == n for r in scores)
        ref = torch.topk(scores, k, dim=1)
        got = topk.topk_indices(scores, k)
        assert got.dtype == torch.int64 and torch.equal(got, ref.indices)
        tied = torch.round(scores * 2) / 2                                  # ~20 distinct values
        assert torch.equal(torch.gather(tied, 1, topk.topk_indices(tied, k)), torch.topk(tied, k, dim=1).values)
    row = torch.tensor([[0.5, float("nan"), 0.5, -0.0, 0.0, float("inf"), 0.5, float("-inf")]])
    assert topk.topk_indices(row, 8).tolist() == [[1, 5, 0, 2, 6, 3, 4, 7]]
