"""GPU: Kernel 3 fused (csrc/locw_tc.cu: both query projections of MSDeformAttn + softmax + location arithmetic in one
tcgen05 kernel, bf16 operands) against (i) an fp64 evaluation of transformer.py:278-293 on the same bf16-rounded
operands and (ii) the unfused CUDA path (library GEMM + tamtr_locw_forward).  Both accumulate the bf16 products in
fp32, so they agree far inside the north star's bf16 tolerance (2e-2); the bound used here is 2e-5 / 1e-4."""
import pytest
import torch

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu


def _inputs(seed, M, C, H, L, P):
    S = L * P
    q = seeding.seeded_tensor(seed, "q", (1, M, C)).bfloat16()
    w_off = (seeding.seeded_tensor(seed, "w_off", (H * S * 2, C)) * 0.05).bfloat16()
    w_att = (seeding.seeded_tensor(seed, "w_att", (H * S, C)) * 0.05).bfloat16()
    b_off = seeding.seeded_tensor(seed, "b_off", (H * S * 2,)) * 2.0
    b_att = seeding.seeded_tensor(seed, "b_att", (H * S,))
    ref = torch.cat([seeding.seeded_uniform(seed, "xy", (1, M, 1, 2)), seeding.seeded_uniform(seed, "wh", (1, M, 1, 2), 0.01, 0.3)], -1)
    return q, ref, w_off, b_off, w_att, b_att          # argument order of ops.sampling_locations_and_weights


def _fp64(q, ref, w_off, b_off, w_att, b_att, H, L, P):
    """transformer.py:278-293 in fp64 on the given (bf16-rounded) operands."""
    M = q.shape[1]
    off = (q.double() @ w_off.double().t() + b_off.double()).view(1, M, H, L, P, 2)
    lg = (q.double() @ w_att.double().t() + b_att.double()).view(1, M, H, L * P)
    attn = torch.softmax(lg, -1).view(1, M, H, L, P)
    r = ref.double()
    loc = r[:, :, None, :, None, :2] + off / P * r[:, :, None, :, None, 2:] * 0.5
    return loc, attn


@pytest.mark.parametrize("M,C,H,L,P", [(4800, 512, 8, 3, 4), (300, 256, 8, 3, 4), (77, 512, 8, 3, 4), (129, 256, 8, 4, 4),
                                       (1000, 128, 4, 3, 4), (200, 512, 16, 4, 4)])
def test_fused_projection_matches_fp64_and_unfused(cuda_lib, M, C, H, L, P):
    ops = cuda_lib.ops
    assert cuda_lib._lib.lib().tamtr_locw_tc_supported(M, C, H, L, P, 1, 4) == 1
    ins = _inputs(M + C, M, C, H, L, P)
    shapes = [[8, 8]] * L
    dev = [t.cuda() for t in ins]
    before = cuda_lib.launch_count()
    loc, attn = ops.sampling_locations_and_weights(*dev, shapes, H, L, P)
    assert loc.dtype == torch.float32 and loc.shape == (1, M, H, L, P, 2) and attn.shape == (1, M, H, L, P)
    assert cuda_lib.launch_count() - before == 1                      # ONE launch of ours, no separate epilogue
    loc64, attn64 = _fp64(*ins, H, L, P)
    assert rel_l2(loc, loc64) < 2e-5 and rel_l2(attn, attn64) < 2e-5
    assert (attn.sum((-1, -2)) - 1).abs().max() < 1e-5
    ops.FUSED_PROJECTION = False
    try:
        loc_u, attn_u = ops.sampling_locations_and_weights(*dev, shapes, H, L, P)
    finally:
        ops.FUSED_PROJECTION = True
    assert rel_l2(loc, loc_u) < 2e-5 and rel_l2(attn, attn_u) < 2e-5


def test_fused_projection_backward_equals_unfused(cuda_lib):
    """The backward is shared (tamtr_locw_backward + two GEMMs); with the fused forward `raw` is materialised only when the
    reference boxes need a gradient.  Both cases against the unfused path."""
    ops = cuda_lib.ops
    M, C, H, L, P = 600, 512, 8, 3, 4
    ins = _inputs(5, M, C, H, L, P)
    shapes = [[8, 8]] * L
    g_loc = seeding.seeded_tensor(6, "g_loc", (1, M, H, L, P, 2)).cuda()
    g_att = seeding.seeded_tensor(6, "g_att", (1, M, H, L, P)).cuda()
    for ref_grad in (False, True):
        grads = []
        for fused in (True, False):
            ops.FUSED_PROJECTION = fused
            try:
                leaf = [ins[0].cuda().requires_grad_(), ins[1].cuda().requires_grad_(ref_grad)] + \
                    [t.cuda().requires_grad_() for t in ins[2:]]
                loc, attn = ops.sampling_locations_and_weights(*leaf, shapes, H, L, P)
                ((loc * g_loc).sum() + (attn * g_att).sum()).backward()
                grads.append([t.grad for t in leaf])
            finally:
                ops.FUSED_PROJECTION = True
        for a, b in zip(*grads):
            assert (a is None) == (b is None)
            if a is not None:
                assert rel_l2(a, b) < 2e-3, rel_l2(a, b)              # bf16 gradients of bf16 leaves: one rounding apart
        assert (grads[0][1] is not None) == ref_grad


def test_unsupported_shapes_take_the_unfused_path(cuda_lib):
    lib = cuda_lib._lib.lib()
    assert lib.tamtr_locw_tc_supported(100, 512, 8, 3, 4, 3, 2) == 0     # 2-d reference points per level
    assert lib.tamtr_locw_tc_supported(100, 500, 8, 3, 4, 1, 4) == 0     # C not a multiple of 64
    assert lib.tamtr_locw_tc_supported(100, 512, 8, 3, 3, 1, 4) == 0     # 9 samples
    assert lib.tamtr_locw_tc_supported(100, 512, 16, 4, 4, 1, 4) == 1    # 768 columns: the heads are split over CTAs
    ops = cuda_lib.ops
    q, _, w_off, b_off, w_att, b_att = _inputs(9, 50, 512, 8, 3, 4)
    ref2 = seeding.seeded_uniform(9, "ref2", (1, 50, 3, 2))
    loc, attn = ops.sampling_locations_and_weights(q.cuda(), ref2.cuda(), w_off.cuda(), b_off.cuda(), w_att.cuda(), b_att.cuda(),
                                                   [[8, 8], [4, 4], [2, 2]], 8, 3, 4)
    assert loc.shape == (1, 50, 8, 3, 4, 2) and torch.isfinite(loc).all()
