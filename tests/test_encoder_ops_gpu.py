"""GPU unit tests of the encoder-side ops (csrc/encoder.cu, box refine, sparse-gradient plumbing) against plain
PyTorch restatements of the reference lines they replace (run on the CPU in fp32/fp64)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from helpers import rel_l2
from oracle import head_ref, seeding

pytestmark = pytest.mark.gpu


def _projs(ch, d, seed):
    projs = nn.ModuleList(nn.Sequential(nn.Conv2d(c, d, 1, bias=False), nn.BatchNorm2d(d)) for c in ch)
    seeding.seeded_fill(projs, seed)
    return projs


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_input_proj_tokens_matches_conv_bn_permute_cat(cuda_lib, training, dtype):
    """head.py:1202-1218: per level conv1x1 + BatchNorm2d, flatten(2).permute(0,2,1), cat."""
    from tamtr_b200 import ops
    import copy
    ch, d, B, sizes = (16, 32, 64), 64, 3, ((12, 10), (6, 5), (3, 3))
    ref = _projs(ch, d, 9).train(training)
    ours = copy.deepcopy(ref).cuda().train(training)
    xs = [seeding.seeded_tensor(8, f"x{i}", (B, c, h, w)) for i, (c, (h, w)) in enumerate(zip(ch, sizes))]
    if dtype == torch.bfloat16:
        xs = [x.bfloat16().float() for x in xs]
        with torch.no_grad():
            for p in ref:
                p[0].weight.copy_(p[0].weight.bfloat16().float())
                ours_p = ours[list(ref).index(p)]
                ours_p[0].weight.copy_(p[0].weight.cuda())
    xr = [x.clone().requires_grad_() for x in xs]
    feats_r = torch.cat([p(x).flatten(2).permute(0, 2, 1) for p, x in zip(ref, xr)], 1)
    probe = seeding.seeded_tensor(8, "probe", feats_r.shape)
    (feats_r * probe).sum().backward()
    xc = [x.cuda().to(dtype).requires_grad_() for x in xs]
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
        feats, shapes = ops.input_proj_tokens(xc, ours, training)
    assert shapes == [list(s) for s in sizes] and feats.dtype == dtype
    (feats.float() * probe.cuda()).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert rel_l2(feats, feats_r) < tol
    for a, b in zip(xc, xr):
        assert rel_l2(a.grad, b.grad) < tol
    for po, pr in zip(ours, ref):
        assert rel_l2(po[0].weight.grad, pr[0].weight.grad) < tol
        assert rel_l2(po[1].weight.grad, pr[1].weight.grad) < tol and rel_l2(po[1].bias.grad, pr[1].bias.grad) < tol
        assert rel_l2(po[1].running_mean, pr[1].running_mean) < tol and rel_l2(po[1].running_var, pr[1].running_var) < tol
        assert int(po[1].num_batches_tracked) == int(pr[1].num_batches_tracked)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_rank_tokens_matches_dense_ranking(cuda_lib, dtype):
    """head.py:1229-1237: max over classes of enc_score_head(LayerNorm(Linear(valid * feats)))."""
    from tamtr_b200 import ops
    B, Lv, d, nc = 3, 777, 512, 10
    lin, ln, sc = nn.Linear(d, d), nn.LayerNorm(d), nn.Linear(d, nc)
    for i, m in enumerate((lin, ln, sc)):
        seeding.seeded_fill(m, 20 + i)
    feats = seeding.seeded_tensor(21, "f", (B, Lv, d))
    valid = (seeding.seeded_uniform(21, "v", (1, Lv, 1)) > 0.1)
    if dtype == torch.bfloat16:
        feats = feats.bfloat16().float()
    ref = sc(ln(lin(valid * feats))).max(-1).values
    # called under autocast for bf16, as the heads do (an outer autocast must not leak into the fp32 kernel operands)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(dtype == torch.bfloat16)):
        out = ops.rank_tokens(feats.cuda().to(dtype), valid.view(-1).to(torch.uint8).cuda(), lin.cuda(), ln.cuda(), sc.cuda())
    assert out.shape == (B, Lv) and out.dtype == torch.float32
    if dtype == torch.float32:
        assert (out.cpu() - ref).abs().max() < 1e-4 * ref.abs().max()
        assert torch.equal(torch.topk(out.cpu(), 50, dim=1).indices.sort(1).values,
                           torch.topk(ref, 50, dim=1).indices.sort(1).values)
    else:
        assert rel_l2(out, ref) < 2e-2
    # nc > 32 takes the strided loop
    sc2 = nn.Linear(d, 80)
    seeding.seeded_fill(sc2, 30)
    ref2 = sc2(ln.cpu()(lin.cpu()(valid * feats))).max(-1).values
    out2 = ops.rank_tokens(feats.cuda().to(dtype), valid.view(-1).to(torch.uint8).cuda(), lin.cuda(), ln.cuda(), sc2.cuda())
    assert rel_l2(out2, ref2) < (1e-4 if dtype == torch.float32 else 2e-2)


def test_box_refine_matches_reference_formula(cuda_lib):
    """transformer.py:875: sigmoid(bbox + inverse_sigmoid(ref)), utils.py:34-39 clamps included."""
    from tamtr_b200 import ops
    bbox = seeding.seeded_tensor(1, "b", (4, 50, 4))
    ref = seeding.seeded_uniform(1, "r", (4, 50, 4), -0.1, 1.1)          # also outside [0, 1] and inside the eps bands
    ref[0, 0] = torch.tensor([0.0, 1.0, 1e-6, 1 - 1e-6])
    br, rr = bbox.clone().requires_grad_(), ref.clone().requires_grad_()
    y = torch.sigmoid(br + head_ref.inverse_sigmoid(rr))
    probe = seeding.seeded_tensor(1, "p", y.shape)
    (y * probe).sum().backward()
    bc, rc = bbox.cuda().requires_grad_(), ref.cuda().requires_grad_()
    yc = ops.box_refine(bc, rc)
    (yc * probe.cuda()).sum().backward()
    assert rel_l2(yc, y) < 1e-6 and rel_l2(bc.grad, br.grad) < 1e-5 and rel_l2(rc.grad, rr.grad) < 1e-5


def test_row_sparse_gradient_hub(cuda_lib):
    from tamtr_b200 import ops
    x = seeding.seeded_tensor(2, "x", (2, 30, 16)).cuda().requires_grad_()
    idx = torch.tensor([3, 3, 17, 45, 59]).cuda()
    hub = ops.GradHub()
    h = ops.grad_hub(x * 1.0, hub)
    rows = ops.select_rows(h, idx, hub)
    ((rows * 2.0).sum() + (h * h).sum()).backward()
    xr = x.detach().clone().requires_grad_()
    hr = xr * 1.0
    ((hr.reshape(-1, 16)[idx] * 2.0).sum() + (hr * hr).sum()).backward()
    assert torch.allclose(x.grad, xr.grad, atol=1e-6) and not hub.pending
