"""GPU: the ragged-point sampler (tamtr_msda_*_ragged through the C ABI) and the modules built on it -- MSDeformAttncls /
MSDeformAttnbox / DecouplingDeformableTransformerDecoderLayer -- against what the reference produced
(utils.py:92-191, transformer.py:300-495, 561-658; tests/golden/msda_ragged.pt) and against the C oracle.
fp32 <= 1e-4 relative, bf16 <= 2e-2 relative on bf16-rounded inputs, corner indices bit-exact."""
import pytest
import torch

from helpers import check_full_or_subset, filled_state_dict, load_golden, probe_loss, rel_l2, subset_err
from oracle import msda, seeding

pytestmark = pytest.mark.gpu

FP32_TOL, BF16_TOL = 1e-4, 2e-2
CASES = [k + "_" + n for k in ("cls", "box") for n in ("tiny_nonsquare", "small_dh32", "small_dh64")]


@pytest.fixture(scope="module")
def ragged():
    return load_golden("msda_ragged")


def _run(fn, value, shapes, loc, attn, grad_out, dtype):
    v = value.cuda().to(dtype).requires_grad_()
    l, a = loc.cuda().requires_grad_(), attn.cuda().requires_grad_()
    out = fn(v, shapes, l, a)
    out.backward(grad_out.cuda().to(dtype))
    return out.detach(), v.grad, l.grad, a.grad


@pytest.mark.parametrize("name", CASES)
def test_fp32_matches_reference_golden_and_oracle(cuda_lib, ragged, name):
    c = ragged["cases"][name]
    fn = cuda_lib.ops.ms_deform_attn_cls if name.startswith("cls") else cuda_lib.ops.ms_deform_attn_box
    value, loc, attn, grad_out = msda.make_ragged_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"], c["points"])
    out, gv, gl, ga = _run(fn, value, c["shapes"], loc, attn, grad_out, torch.float32)
    assert out.shape == c["out"].shape and gl.shape == c["grad_loc"].shape and ga.shape == c["grad_attn"].shape
    assert rel_l2(out, c["out"]) < FP32_TOL
    check_full_or_subset(gv, c, "grad_value", FP32_TOL)
    assert rel_l2(gl, c["grad_loc"]) < FP32_TOL and rel_l2(ga, c["grad_attn"]) < FP32_TOL
    o_ref = msda.forward_ragged_c(value, c["shapes"], loc, attn, c["points"])
    gv_ref, gl_ref, ga_ref = msda.backward_ragged_c(grad_out, value, c["shapes"], loc, attn, c["points"])
    assert (out.cpu() - o_ref).abs().max() < 1e-4 * o_ref.abs().max()
    assert rel_l2(gv, gv_ref) < FP32_TOL and rel_l2(gl, gl_ref) < FP32_TOL
    assert rel_l2(ga.reshape(ga_ref.shape), ga_ref) < FP32_TOL


@pytest.mark.parametrize("name", ["cls_small_dh32", "box_small_dh64"])
def test_bf16_on_bf16_rounded_inputs(cuda_lib, ragged, name):
    c = ragged["cases"][name]
    fn = cuda_lib.ops.ms_deform_attn_cls if name.startswith("cls") else cuda_lib.ops.ms_deform_attn_box
    value, loc, attn, grad_out = msda.make_ragged_inputs(c["seed"], c["B"], c["Lq"], c["H"], c["Dh"], c["shapes"], c["points"])
    v_r, g_r = value.bfloat16().float(), grad_out.bfloat16().float()
    out, gv, gl, ga = _run(fn, v_r, c["shapes"], loc, attn, g_r, torch.bfloat16)
    assert out.dtype == torch.bfloat16 and gv.dtype == torch.bfloat16 and gl.dtype == torch.float32
    o_ref = msda.forward_ragged_c(v_r, c["shapes"], loc, attn, c["points"])
    gv_ref, gl_ref, ga_ref = msda.backward_ragged_c(g_r, v_r, c["shapes"], loc, attn, c["points"])
    assert rel_l2(out, o_ref) < BF16_TOL and rel_l2(gv, gv_ref) < BF16_TOL
    assert rel_l2(gl, gl_ref) < BF16_TOL and rel_l2(ga.reshape(ga_ref.shape), ga_ref) < BF16_TOL


@pytest.mark.parametrize("points", [(2, 4, 6), (6, 4, 2), (1, 1, 30), (5, 3, 1)])
def test_general_point_counts_and_corner_indices(cuda_lib, points):
    """Any split with sum <= 32 (not only the reference's two): forward against the oracle, corner indices and in-bounds
    flags bit-exact at exact pixel centres / edges +-2 ulp of every level."""
    shapes = [[20, 24], [10, 12], [5, 6]]
    S = sum(points)
    g = torch.Generator().manual_seed(S)
    B, Lq, H, Dh = 2, 33, 4, 16
    value = torch.randn(B, 20 * 24 + 10 * 12 + 5 * 6, H, Dh, generator=g)
    loc = torch.rand(B, Lq, H, S, 2, generator=g) * 1.4 - 0.2
    attn = torch.softmax(torch.randn(B, Lq, H, S, generator=g), -1)
    out = cuda_lib.ops.ms_deform_attn_ragged(value.cuda(), shapes, loc.cuda(), attn.cuda(), points)
    assert rel_l2(out, msda.forward_ragged_c(value, shapes, loc, attn, points)) < FP32_TOL
    adv = msda.adversarial_locations(shapes, H=4, P=4)[:, :, :, :, 0]            # [1, Lq, H, L, 2]: one probe per level
    adv_loc = torch.cat([adv[:, :, :, l:l + 1].expand(-1, -1, -1, p, -1) for l, p in enumerate(points)], 3).contiguous()
    for probe in (adv_loc, loc):
        x0, y0, inb = cuda_lib.ops.ms_deform_attn_corners_ragged(probe.cuda(), shapes, points)
        rx0, ry0, rinb = msda.corners_ragged_c(probe, shapes, points)
        assert torch.equal(inb.cpu(), rinb)
        live = rinb.bool().any(-1)
        assert torch.equal(x0.cpu()[live], rx0[live]) and torch.equal(y0.cpu()[live], ry0[live])


def test_uniform_split_equals_the_base_sampler(cuda_lib):
    """points = (4, 4, 4) is the base op: bit-identical output (same kernel, same tap order)."""
    shapes = [[16, 16], [8, 8], [4, 4]]
    value, loc, attn, _ = msda.make_inputs(7, 2, 40, 8, 32, shapes, oob_frac=0.2)
    a = cuda_lib.ops.ms_deform_attn(value.cuda(), shapes, loc.cuda(), attn.cuda())
    b = cuda_lib.ops.ms_deform_attn_ragged(value.cuda(), shapes, loc.view(2, 40, 8, 12, 2).cuda(), attn.cuda(), (4, 4, 4))
    assert torch.equal(a, b)


def test_bad_arguments_raise(cuda_lib):
    shapes = [[4, 4], [2, 2], [1, 1]]
    value = torch.zeros(1, 21, 2, 8, device="cuda")
    loc = torch.zeros(1, 3, 2, 12, 2, device="cuda")
    attn = torch.zeros(1, 3, 2, 3, 4, device="cuda")
    with pytest.raises(RuntimeError, match="point counts"):
        cuda_lib.ops.ms_deform_attn_ragged(value, shapes, loc, attn, (6, 6))
    with pytest.raises(RuntimeError, match="sampling_locations"):
        cuda_lib.ops.ms_deform_attn_ragged(value, shapes, loc, attn, (2, 4, 7))
    with pytest.raises(RuntimeError, match="is_cuda"):
        cuda_lib.ops.ms_deform_attn_cls(value.cpu(), shapes, loc.cpu(), attn.cpu())


def test_decoupling_decoder_layer_fp32(cuda_lib, ragged):
    from tamtr_b200.modules import DecouplingDeformableTransformerDecoderLayer
    c = ragged["layer"]
    m = DecouplingDeformableTransformerDecoderLayer(c["d"], c["H"], c["d_ffn"], 0.0, torch.nn.ReLU(), 3, 4)
    filled_state_dict(m, c["fill_seed"], c["manifest"])          # same state_dict keys / shapes as the reference's class
    m.cuda()
    B, Lq, d = c["B"], c["Lq"], c["d"]
    embed = seeding.seeded_tensor(45, "embed", (B, Lq, d)).cuda().requires_grad_()
    embed1 = seeding.seeded_tensor(45, "embed1", (B, Lq, d)).cuda().requires_grad_()
    feats = seeding.seeded_smooth_tokens(45, "feats", B, d, c["shapes"], factor=2).cuda().requires_grad_()
    ref_box = torch.cat([seeding.seeded_uniform(45, "ref_xy", (B, Lq, 2)),
                         seeding.seeded_uniform(45, "ref_wh", (B, Lq, 2), 0.01, 0.3)], -1).cuda()
    pos = seeding.seeded_tensor(45, "pos", (B, Lq, d)).cuda()
    mask = torch.zeros(Lq, Lq, dtype=torch.bool)
    mask[:16, 16:] = True
    mask[16:, :16] = True
    o_cls, o_box = m(embed, embed1, ref_box, feats, c["shapes"], None, mask.cuda(), pos)
    (probe_loss(o_cls, 46, "p_cls") + probe_loss(o_box, 46, "p_box")).backward()
    assert rel_l2(o_cls, c["out_cls"]) < FP32_TOL and rel_l2(o_box, c["out_box"]) < FP32_TOL
    assert rel_l2(embed.grad, c["grad_embed"]) < FP32_TOL and rel_l2(embed1.grad, c["grad_embed1"]) < FP32_TOL
    assert subset_err(feats.grad, c["grad_feats_subset"]) < FP32_TOL
    for k, p in m.named_parameters():
        g = c["param_grads"].get(k)
        if g is None:
            assert p.grad is None or not p.grad.any(), k
        elif isinstance(g, tuple):
            assert subset_err(p.grad, g) < 2e-4, k
        else:
            assert rel_l2(p.grad, g) < 2e-4, (k, rel_l2(p.grad, g))
