"""CPU: the ragged-point restatements (oracle/msda.py: forward_ragged_c / backward_ragged_c on the C core, and the
grid_sample op sequence) pinned to what the reference's multi_scale_deformable_attn_pytorch_cls / _box produced
(ultralytics/nn/modules/utils.py:92-191; tests/golden/msda_ragged.pt from oracle/make_goldens.py ragged)."""
import pytest
import torch

from helpers import check_full_or_subset, load_golden, rel_l2
from oracle import msda

CASES = [k + "_" + n for k in ("cls", "box") for n in ("tiny_nonsquare", "small_dh32", "small_dh64")]


@pytest.fixture(scope="module")
def ragged():
    return load_golden("msda_ragged")


@pytest.mark.parametrize("name", CASES)
def test_c_oracle_matches_reference_golden(ragged, name):
    c = ragged["cases"][name]
    assert tuple(c["points"]) == ((2, 4, 6) if name.startswith("cls") else (6, 4, 2))
    value, loc, attn, grad_out = msda.make_ragged_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"], c["points"])
    out = msda.forward_ragged_c(value, c["shapes"], loc, attn, c["points"])
    gv, gl, ga = msda.backward_ragged_c(grad_out, value, c["shapes"], loc, attn, c["points"])
    assert rel_l2(out, c["out"]) < 5e-6
    check_full_or_subset(gv, c, "grad_value", 5e-6)
    assert rel_l2(gl, c["grad_loc"]) < 5e-6
    assert rel_l2(ga.reshape(c["grad_attn"].shape), c["grad_attn"]) < 5e-6


@pytest.mark.parametrize("name", ["cls_tiny_nonsquare", "box_tiny_nonsquare"])
def test_gridsample_restatement(ragged, name):
    c = ragged["cases"][name]
    value, loc, attn, grad_out = msda.make_ragged_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"], c["points"])
    v, l, a = (t.clone().requires_grad_() for t in (value, loc, attn))
    out = msda.msda_ragged_gridsample_torch(v, c["shapes"], l, a, c["points"])
    out.backward(grad_out)
    assert torch.equal(out.detach(), c["out"])            # the same library calls in the same order
    assert rel_l2(v.grad, c["grad_value"]) < 1e-6 and rel_l2(l.grad, c["grad_loc"]) < 1e-6
    assert rel_l2(a.grad, c["grad_attn"]) < 1e-6


def test_split_order_matters(ragged):
    """cls and box differ only in which samples go to which level: feeding the cls split to the box data must NOT agree."""
    c = ragged["cases"]["box_small_dh32"]
    value, loc, attn, _ = msda.make_ragged_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"], c["points"])
    wrong = msda.forward_ragged_c(value, c["shapes"], loc, attn, (2, 4, 6))
    assert rel_l2(wrong, c["out"]) > 0.1
