"""GPU parity of the fused residual + LayerNorm kernels (csrc/layernorm.cu) against the reference's op sequence
`norm(embed + tgt)` (ultralytics/nn/modules/transformer.py:548,553,537) evaluated by torch on the CPU in fp64/fp32."""
import pytest
import torch
import torch.nn as nn

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu


def _ref(x, res, w, b, eps, probe):
    x, w, b = x.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
    r = None if res is None else res.double().requires_grad_()
    y = torch.nn.functional.layer_norm(x if r is None else x + r, (x.shape[-1],), w, b, eps)
    (y * probe.double()).sum().backward()
    return y.detach(), x.grad, None if r is None else r.grad, w.grad, b.grad


@pytest.mark.parametrize("rows_shape,d", [((16, 300), 512), ((2, 37), 256), ((1, 1), 128), ((3, 1000), 384)])
@pytest.mark.parametrize("xdt,rdt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.bfloat16), (torch.float32, None)])
def test_add_layer_norm_matches_reference(cuda_lib, rows_shape, d, xdt, rdt):
    from tamtr_b200 import ops
    shape = (*rows_shape, d)
    x = (seeding.seeded_tensor(d, "x", shape) * 2 + 0.5).to(xdt)
    res = None if rdt is None else seeding.seeded_tensor(d, "r", shape).to(rdt)
    norm = nn.LayerNorm(d)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.2 * seeding.seeded_tensor(d, "w", (d,)))
        norm.bias.copy_(0.1 * seeding.seeded_tensor(d, "b", (d,)))
    probe = seeding.seeded_tensor(d, "p", shape)
    y_ref, gx_ref, gr_ref, gw_ref, gb_ref = _ref(x.float(), None if res is None else res.float(), norm.weight.detach(),
                                                 norm.bias.detach(), norm.eps, probe)
    norm = norm.cuda()
    xc = x.cuda().requires_grad_()
    rc = None if res is None else res.cuda().requires_grad_()
    y = ops.add_layer_norm(xc, rc, norm)
    all_bf16 = xdt == torch.bfloat16 and rdt in (torch.bfloat16, None)
    assert y.dtype == (torch.bfloat16 if all_bf16 else torch.float32)
    (y.float() * probe.cuda()).sum().backward()
    out_tol = 5e-3 if all_bf16 else 1e-5
    assert rel_l2(y, y_ref) < out_tol, rel_l2(y, y_ref)
    # gradients w.r.t. bf16 inputs are rounded once to bf16
    assert rel_l2(xc.grad, gx_ref) < (5e-3 if xdt == torch.bfloat16 else 1e-5)
    if rc is not None:
        assert rc.grad.dtype == rdt and rel_l2(rc.grad, gr_ref) < (5e-3 if rdt == torch.bfloat16 else 1e-5)
    par_tol = 5e-3 if all_bf16 else 1e-5             # an all-bf16 output hands the backward a bf16-rounded gradient
    assert rel_l2(norm.weight.grad, gw_ref) < par_tol and rel_l2(norm.bias.grad, gb_ref) < par_tol


def test_decoder_layer_uses_fused_norms(cuda_lib):
    """3 fused forward + 3 fused backward launches per decoder layer."""
    import tamtr_b200
    from tamtr_b200 import _lib
    from tamtr_b200.modules import DeformableTransformerDecoderLayer
    layer = DeformableTransformerDecoderLayer(256, 8, 512, 0.0, nn.ReLU(), 3, 4).cuda()
    shapes = [[20, 20], [10, 10], [5, 5]]
    B, Lq = 2, 50
    embed = torch.randn(B, Lq, 256, device="cuda", requires_grad=True)
    feats = torch.randn(B, 525, 256, device="cuda")
    refer = torch.rand(B, Lq, 4, device="cuda")
    _lib.profile_enable(True)
    out = layer(embed, refer, feats, shapes, None, None, torch.randn(B, Lq, 256, device="cuda"))
    out.sum().backward()
    torch.cuda.synchronize()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    assert prof["add_layernorm_fwd"][1] == 3 and prof["add_layernorm_bwd"][1] == 3


@pytest.mark.parametrize("rows_shape,d", [((16, 290), 512), ((2, 37), 256), ((1, 1), 128)])
@pytest.mark.parametrize("pdt", [torch.bfloat16, torch.float32, None])
@pytest.mark.parametrize("want_lp", [True, False])
def test_add_layer_norm_side_outputs(cuda_lib, rows_shape, d, pdt, want_lp):
    """norm(x + res) with the bf16 side outputs against the op sequence autocast runs (transformer.py:548-552:
    embed = norm(embed + tgt); query = embed + pos, both rounded to bf16 in front of the projections): the forward values are
    the same roundings of the same fp32 numbers (bit-equal), the backward sums the three output gradients in fp32."""
    from tamtr_b200 import ops
    shape = (*rows_shape, d)
    x = (seeding.seeded_tensor(d, "x", shape) * 2 + 0.5).cuda()
    res = seeding.seeded_tensor(d, "r", shape).bfloat16().cuda()
    pos = None if pdt is None else seeding.seeded_tensor(d, "pos", shape).to(pdt).cuda()
    norm = nn.LayerNorm(d)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.2 * seeding.seeded_tensor(d, "w", (d,)))
        norm.bias.copy_(0.1 * seeding.seeded_tensor(d, "b", (d,)))
    norm = norm.cuda()
    probes = [seeding.seeded_tensor(d, f"p{i}", shape).cuda() for i in range(3)]

    def run(fused):
        xs, rs = x.clone().requires_grad_(), res.clone().requires_grad_()
        ps = None if pos is None else pos.clone().requires_grad_()
        norm.zero_grad(set_to_none=True)
        if fused:
            y, y_lp, q_lp = ops.add_layer_norm_sides(xs, rs, norm, pos=ps, want_lp=want_lp)
        else:
            y = ops.add_layer_norm(xs, rs, norm)
            y_lp = y.to(torch.bfloat16) if want_lp else None
            q_lp = None if ps is None else (y + ps).to(torch.bfloat16)
        loss = (y * probes[0]).sum()
        if y_lp is not None:
            loss = loss + (y_lp * probes[1].bfloat16()).float().sum()
        if q_lp is not None:
            loss = loss + (q_lp * probes[2].bfloat16()).float().sum()
        loss.backward()
        return (y, y_lp, q_lp), (xs.grad, rs.grad, None if ps is None else ps.grad, norm.weight.grad.clone(),
                                 norm.bias.grad.clone())

    outs_f, grads_f = run(True)
    outs_u, grads_u = run(False)
    for a, b in zip(outs_f, outs_u):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.dtype == b.dtype and torch.equal(a, b)
    for a, b in zip(grads_f, grads_u):
        assert (a is None) == (b is None)
        if a is not None:
            assert a.dtype == b.dtype and rel_l2(a.float(), b.float()) < 5e-3


def test_pos_cast_and_its_gradient(cuda_lib):
    from tamtr_b200 import ops
    shape = (3, 41, 256)
    x = seeding.seeded_tensor(1, "x", shape).cuda().requires_grad_()
    pos = seeding.seeded_tensor(1, "pos", shape).bfloat16().cuda().requires_grad_()
    p = [seeding.seeded_tensor(1, f"p{i}", shape).cuda() for i in range(3)]
    x0, x_lp, q_lp = ops.pos_cast(x, pos)
    assert torch.equal(x0, x) and torch.equal(x_lp, x.to(torch.bfloat16)) and torch.equal(q_lp, (x + pos).to(torch.bfloat16))
    ((x0 * p[0]).sum() + (x_lp * p[1].bfloat16()).float().sum() + (q_lp * p[2].bfloat16()).float().sum()).backward()
    gx, gp = x.grad.clone(), pos.grad.clone()
    x.grad = pos.grad = None
    ((x * p[0]).sum() + (x.to(torch.bfloat16) * p[1].bfloat16()).float().sum()
     + ((x + pos).to(torch.bfloat16) * p[2].bfloat16()).float().sum()).backward()
    assert gx.dtype == x.grad.dtype and rel_l2(gx, x.grad) < 1e-6
    assert gp.dtype == pos.grad.dtype and torch.equal(gp, pos.grad)
    # a single consumer: the gradient passes through
    x.grad = None
    x0, x_lp, q_lp = ops.pos_cast(x, pos)
    (x_lp.float() * p[1]).sum().backward()
    assert rel_l2(x.grad, p[1].bfloat16().float()) < 1e-6


@pytest.mark.parametrize("train", [True, False])
def test_decoder_layer_low_precision_operands_from_the_norm_kernels(cuda_lib, monkeypatch, train):
    """Under bf16 autocast the layer takes its bf16 operands from the add + LayerNorm / pos_cast kernels: same forward
    values as the autocast op sequence (bit-equal), gradients equal up to the order of the fp32 sums, and no ATen cast or
    add launch of the stream's size left in the forward."""
    from tamtr_b200 import modules
    from tamtr_b200.modules import DeformableTransformerDecoderLayer
    torch.manual_seed(0)
    layer = DeformableTransformerDecoderLayer(256, 8, 512, 0.0, nn.ReLU(), 3, 4).cuda().train(train)
    seeding.seeded_fill(layer, 5)
    shapes = [[20, 20], [10, 10], [5, 5]]
    B, Lq = 2, 50
    embed = seeding.seeded_tensor(2, "e", (B, Lq, 256)).cuda()
    feats = seeding.seeded_tensor(2, "f", (B, 525, 256)).cuda()
    refer = seeding.seeded_tensor(2, "r", (B, Lq, 4)).sigmoid().cuda()
    pos = seeding.seeded_tensor(2, "pos", (B, Lq, 256)).bfloat16().cuda()
    probe = seeding.seeded_tensor(2, "probe", (B, Lq, 256)).cuda()

    def run(lowp):
        monkeypatch.setattr(modules, "LOWP_LAYER", lowp)
        e, p = embed.clone().requires_grad_(), pos.clone().requires_grad_()
        layer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = layer(e, refer, feats, shapes, None, None, p)
        (out.float() * probe).sum().backward()
        grads = {n: q.grad.float().clone() for n, q in layer.named_parameters() if q.grad is not None}
        grads["embed"], grads["pos"] = e.grad.clone(), p.grad.float().clone()
        return out.detach(), grads

    out_l, g_l = run(True)
    out_u, g_u = run(False)
    assert out_l.dtype == out_u.dtype and torch.equal(out_l, out_u)
    assert set(g_l) == set(g_u)
    for k in g_u:
        assert rel_l2(g_l[k], g_u[k]) < 1e-2, k
    # launch census of the forward
    from torch.profiler import profile, ProfilerActivity
    counts = {}
    for lowp in (True, False):
        monkeypatch.setattr(modules, "LOWP_LAYER", lowp)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            layer(embed, refer, feats, shapes, None, None, pos)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                layer(embed, refer, feats, shapes, None, None, pos)
                torch.cuda.synchronize()
        counts[lowp] = sum(1 for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA
                           and "elementwise_kernel" in ev.name)
    assert counts[True] + 5 <= counts[False], counts


# ---------------------------------------------------------------------------------------------- Linear / self-attention
@pytest.mark.parametrize("rows,n,dt", [(4800, 512, torch.bfloat16), (4800, 1024, torch.float32), (37, 8, torch.bfloat16),
                                       (1, 1536, torch.float32), (1000, 136, torch.bfloat16)])
def test_col_sum(cuda_lib, rows, n, dt):
    from tamtr_b200 import ops
    g = seeding.seeded_tensor(rows, "g", (rows, n)).to(dt)
    got = ops.col_sum(g.cuda())
    assert got.dtype == torch.float32
    assert rel_l2(got, g.double().sum(0)) < 1e-6


def test_linear_and_self_attention_match_the_modules(cuda_lib):
    """ops.linear / ops.self_attention against nn.Linear / nn.MultiheadAttention called the way the reference calls
    them (sequence-first through transposes, bool mask with True = blocked; transformer.py:544-547), on the CPU."""
    from tamtr_b200 import ops
    torch.manual_seed(0)
    B, L, d, H = 3, 50, 256, 8
    mha = nn.MultiheadAttention(d, H)
    lin = nn.Linear(d, 72)
    x = seeding.seeded_tensor(1, "x", (B, L, d))
    pos = seeding.seeded_tensor(1, "pos", (B, L, d))
    mask = torch.rand(L, L) < 0.2
    mask.fill_diagonal_(False)
    probe = seeding.seeded_tensor(1, "p", (B, L, 72))

    def run(mod_mha, mod_lin, xx, pp, mm, pr, ours):
        xx = xx.clone().requires_grad_()
        q = xx + pp
        if ours:
            t = ops.self_attention(mod_mha, q, xx, mm)
            y = ops.linear(t, mod_lin)
        else:
            t = mod_mha(q.transpose(0, 1), q.transpose(0, 1), xx.transpose(0, 1), attn_mask=mm)[0].transpose(0, 1)
            y = mod_lin(t)
        (y * pr).sum().backward()
        return y.detach(), xx.grad, [p.grad.clone() for p in list(mod_mha.parameters()) + list(mod_lin.parameters())]

    y_ref, gx_ref, gp_ref = run(mha, lin, x, pos, mask, probe, False)
    mha.zero_grad()
    lin.zero_grad()
    mha.cuda()
    lin.cuda()
    y, gx, gp = run(mha, lin, x.cuda(), pos.cuda(), mask.cuda(), probe.cuda(), True)
    assert rel_l2(y, y_ref) < 1e-5 and rel_l2(gx, gx_ref) < 1e-5
    for a, b in zip(gp, gp_ref):
        assert rel_l2(a, b) < 1e-5
