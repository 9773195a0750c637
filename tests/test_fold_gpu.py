"""GPU tests of the folded encoder side: the two tcgen05 kernels of csrc/tokgemm.cu against fp32 restatements on the same
bf16 operands, and the folded head against the unfolded kernels (same module, `folded_projection` off)."""
import pytest
import torch

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu


def _maps(B, C, H, W, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, C, H, W, generator=g) * 0.7 + 0.3).bfloat16().cuda()


@pytest.mark.parametrize("B,C,H,W,N0,N1,NT", [
    (2, 128, 20, 20, 256, 128, 16),        # two 128-token blocks per CTA, ragged last block (400 tokens)
    (1, 64, 8, 8, 64, 64, 16),             # a single, partly filled block
    (3, 256, 16, 24, 384, 256, 80),        # tail wider than one 64-column chunk (nc = 80 classes)
    (2, 512, 10, 20, 1536, 512, 16),       # one block per CTA (C = 512), TAM-TR column counts
    (2, 128, 40, 40, 1536, 512, 16),       # 1600 tokens: 12.5 blocks per image
    (2, 192, 12, 12, 128, 64, 0),          # no tail
])
def test_tok_project_matches_fp32(cuda_lib, B, C, H, W, N0, N1, NT):
    from tamtr_b200 import fold
    x = _maps(B, C, H, W, 1)
    g = torch.Generator().manual_seed(2)
    Nall = N0 + N1 + NT
    w = (torch.randn(Nall, C, generator=g) / C ** 0.5).bfloat16().cuda()
    bias = torch.randn(Nall, generator=g).cuda()
    HW, pre, post = H * W, 24, 8                       # the level sits inside a longer token axis
    Lv = pre + HW + post
    out0 = torch.full((B, Lv, N0), 7.0, dtype=torch.bfloat16, device="cuda")
    out1 = torch.full((B, Lv, max(N1, 1)), 7.0, dtype=torch.bfloat16, device="cuda")[..., :N1].contiguous() if N1 else None
    raw = torch.full((B, Lv, max(NT, 1)), 7.0, device="cuda")[..., :NT].contiguous() if NT else None
    fold._kernel_project(x, w, bias, out0, out1 if N1 else out0, raw if NT else out0, pre, N0, N1, NT)
    torch.cuda.synchronize()
    ref = x.float().flatten(2).transpose(1, 2) @ w.float().t() + bias                 # [B, HW, Nall]
    assert rel_l2(out0[:, pre:pre + HW], ref[..., :N0]) < 4e-3
    assert (out0[:, pre:pre + HW].float() - ref[..., :N0]).abs().max() < 0.05
    if N1:
        assert rel_l2(out1[:, pre:pre + HW], ref[..., N0:N0 + N1]) < 4e-3
    if NT:
        assert rel_l2(raw[:, pre:pre + HW], ref[..., N0 + N1:]) < 1e-5
    # nothing outside the level's token range is touched
    for t in (out0, out1, raw):
        if t is not None:
            assert torch.all(t[:, :pre] == 7.0) and torch.all(t[:, pre + HW:] == 7.0)


@pytest.mark.parametrize("B,C,H,W,N0,N1,nc", [
    (2, 128, 20, 20, 256, 128, 10),        # two token blocks per CTA: a thread owns its token's whole row
    (2, 512, 10, 20, 1536, 512, 10),       # one block per CTA: two threads share a token (partial moments through smem)
    (2, 128, 40, 40, 1536, 512, 10),       # TAM-TR column counts, 12.5 blocks per image
    (1, 256, 8, 16, 128, 192, 40),         # tail starts in the second chunk of its step; 40 classes
])
def test_tok_project_rank_matches_fp32(cuda_lib, B, C, H, W, N0, N1, nc):
    """ranking scores out of the projection's epilogue == the same formulas in fp32 on the same bf16 operands"""
    from tamtr_b200 import fold
    x = _maps(B, C, H, W, 7)
    g = torch.Generator().manual_seed(8)
    NT = (nc + 1 + 15) // 16 * 16
    Nall = N0 + N1 + NT
    w = (torch.randn(Nall, C, generator=g) / C ** 0.5).bfloat16().cuda()
    bias = torch.randn(Nall, generator=g).cuda()
    consts = torch.randn(2 + 3 * NT, generator=g)
    consts[1] = consts[1].abs() + 40.0                 # sum enc_bias^2: keeps the variance positive
    consts = consts.cuda()
    HW, pre, post = H * W, 24, 8
    Lv = pre + HW + post
    valid = (torch.rand(Lv, generator=g) > 0.25).to(torch.uint8).cuda()
    out0 = torch.full((B, Lv, N0), 7.0, dtype=torch.bfloat16, device="cuda")
    zero = torch.full((B, Lv, N0), 7.0, dtype=torch.bfloat16, device="cuda")      # the gradient arena zeroed alongside
    scores = torch.full((B, Lv), 7.0, device="cuda")
    fold._kernel_project(x, w, bias, out0, None, None, pre, N0, N1, NT, (scores, valid, consts, nc, 1e-5), zero)
    torch.cuda.synchronize()
    assert torch.all(zero[:, pre:pre + HW] == 0) and torch.all(zero[:, :pre] == 7.0) and torch.all(zero[:, pre + HW:] == 7.0)
    y = x.float().flatten(2).transpose(1, 2) @ w.float().t() + bias
    assert rel_l2(out0[:, pre:pre + HW], y[..., :N0]) < 4e-3
    ok = valid[pre:pre + HW].bool().view(1, HW, 1)
    E, tail = y[..., N0:N0 + N1] * ok, y[..., N0 + N1:] * ok
    mean = (E.sum(-1) + consts[0]) / N1
    var = ((E * E).sum(-1) + 2 * tail[..., NT - 1] + consts[1]) / N1 - mean * mean
    rstd = torch.rsqrt(var.clamp_min(0) + 1e-5)
    bw, sw, ck = consts[2:2 + NT], consts[2 + NT:2 + 2 * NT], consts[2 + 2 * NT:2 + 3 * NT]
    ref = (rstd.unsqueeze(-1) * (tail + bw - mean.unsqueeze(-1) * sw) + ck)[..., :nc].max(-1).values
    got = scores[:, pre:pre + HW]
    assert (got - ref).abs().max() < 2e-3 * ref.abs().max(), (got - ref).abs().max()
    assert torch.all(scores[:, :pre] == 7.0) and torch.all(scores[:, pre + HW:] == 7.0)
    assert torch.all(out0[:, :pre] == 7.0) and torch.all(out0[:, pre + HW:] == 7.0)


@pytest.mark.parametrize("B,C,H,W", [(2, 128, 20, 20), (1, 64, 8, 8), (3, 256, 16, 24), (2, 512, 10, 20), (4, 128, 40, 40)])
def test_tok_reduce_moments(cuda_lib, B, C, H, W):
    from tamtr_b200 import fold
    x = _maps(B, C, H, W, 3)
    G, S1 = fold._kernel_reduce(x, H * W, C * H * W, False, x, C)
    X = x.double().flatten(2)
    assert rel_l2(G, torch.einsum("bct,bdt->cd", X, X)) < 1e-5
    assert rel_l2(S1, X.sum((0, 2))) < 1e-5


@pytest.mark.parametrize("B,C,H,W,M", [(2, 128, 20, 20, 256), (1, 64, 8, 8, 64), (3, 256, 16, 24, 384),
                                       (2, 512, 10, 20, 1536), (2, 128, 40, 40, 1536), (2, 128, 12, 12, 192)])
def test_tok_reduce_weight_gradient(cuda_lib, B, C, H, W, M):
    from tamtr_b200 import fold
    x = _maps(B, C, H, W, 4)
    HW, pre, post = H * W, 16, 40
    Lv = pre + HW + post
    g = torch.Generator().manual_seed(5)
    a = torch.randn(B, Lv, M, generator=g).bfloat16().cuda()
    D, rs = fold._kernel_reduce(a[:, pre:], M, Lv * M, True, x, M)
    A = a[:, pre:pre + HW].double()
    assert rel_l2(D, torch.einsum("btm,bct->mc", A, x.double().flatten(2))) < 1e-5
    assert rel_l2(rs, A.sum((0, 1))) < 1e-5


def _head(B=2, sizes=(48, 24, 12)):
    from tamtr_b200.head import ManbaWorldDecoder
    torch.manual_seed(0)
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False).cuda().train()
    seeding.seeded_fill(m, 11)
    xs = [seeding.seeded_smooth_map(5, f"x{i}", (B, c, s, s)).bfloat16().cuda()
          for i, (c, s) in enumerate(zip((128, 256, 512), sizes))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(5, "t", (B, 10, 512)), dim=-1).cuda()
    m.num_denoising = 0                  # no ground truth in these tests: matching queries only
    return m, xs, text


def test_folded_encoder_matches_unfolded(cuda_lib):
    """values of every layer, ranking scores, selected rows and the BatchNorm side effects, folded vs unfolded kernels."""
    import copy
    m, xs, text = _head()
    ref = copy.deepcopy(m)
    ref.folded_projection = False
    with torch.autocast("cuda", dtype=torch.bfloat16):
        tok, shapes, hub = m._encode(xs)
        assert getattr(tok, "is_folded", False) and hub is None
        feats, shapes2, hub2 = ref._encode(xs)
        assert shapes == shapes2
        vals, _ = ref.decoder._project_values(feats, None, 3)
        for a, b in zip(tok.values, vals):
            assert a.shape == b.shape and rel_l2(a, b) < 1.5e-2
        anchors, valid = m._anchors(shapes, torch.bfloat16, xs[0].device)
        r_f, r_u = m._rank_tokens(tok, valid), ref._rank_tokens(feats, valid)
        v = valid.view(1, -1).expand_as(r_f)
        assert rel_l2(r_f[v], r_u[v]) < 1.5e-2
        idx = torch.randint(0, feats.shape[0] * feats.shape[1], (300,), device="cuda")
        assert rel_l2(tok.rows(idx), feats.reshape(-1, 512)[idx]) < 1e-2
    for p, r in zip(m.input_proj, ref.input_proj):
        assert rel_l2(p[1].running_mean, r[1].running_mean) < 1e-2
        assert rel_l2(p[1].running_var, r[1].running_var) < 1e-2
        assert int(p[1].num_batches_tracked) == int(r[1].num_batches_tracked) == 1


def _step(m, xs, text, seed=3, autocast=True):
    for p in m.parameters():
        p.grad = None
    torch.manual_seed(seed)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        out = m(xs, text)
    if not m.training:
        out = out[1]
    db, ds, eb, es = out[:4]
    loss = db.float().square().mean() + 0.1 * ds.float().sigmoid().mean() + eb.float().square().mean() + 0.1 * es.float().sigmoid().mean()
    loss.backward()
    return loss.detach(), {n: p.grad.detach().float().clone() for n, p in m.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("train", [True, False])
def test_folded_head_step_matches_unfolded(cuda_lib, train):
    """whole head forward + backward: same loss and parameter gradients as the unfolded kernels within bf16 tolerance,
    and the folded path launches the projection / reduction kernels instead of the BatchNorm token passes."""
    import copy
    from tamtr_b200 import _lib
    m, xs, text = _head()
    if not train:
        m.eval()
    ref = copy.deepcopy(m)
    ref.folded_projection = False
    _lib.profile_enable(True)
    loss_f, g_f = _step(m, xs, text)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    assert prof.get("tok_project", (0, 0))[1] == 3
    assert prof.get("tok_reduce", (0, 0))[1] == (6 if train else 3)
    loss_u, g_u = _step(ref, xs, text)
    assert abs(loss_f.item() - loss_u.item()) < 2e-2 * abs(loss_u.item())
    assert set(g_f) == set(g_u)
    flat_f = torch.cat([g_f[k].reshape(-1) for k in sorted(g_f)])
    flat_u = torch.cat([g_u[k].reshape(-1) for k in sorted(g_u)])
    assert rel_l2(flat_f, flat_u) < 5e-2
    # per tensor, both bf16 paths are judged against the SAME fp32 run of the unfolded kernels: the folded path (one
    # rounding less: Y and M are never rounded to bf16) must not be further from it than the unfolded bf16 path is
    ref32 = copy.deepcopy(ref).float()
    _, g_32 = _step(ref32, [x.float() for x in xs], text, autocast=False)
    for k in ("input_proj.0.0.weight", "input_proj.2.1.weight", "input_proj.1.1.bias",
              "decoder.layers.0.cross_attn.value_proj.weight", "decoder.layers.2.cross_attn.value_proj.bias",
              "enc_output.0.weight"):
        e_f, e_u = rel_l2(g_f[k], g_32[k]), rel_l2(g_u[k], g_32[k])
        assert e_f < max(1.25 * e_u, 3e-2), (k, e_f, e_u)


def test_folded_head_inference_matches_unfolded(cuda_lib):
    import copy
    m, xs, text = _head(B=1)
    m.eval()
    ref = copy.deepcopy(m)
    ref.folded_projection = False
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y_f, _ = m(xs, text)
        y_u, _ = ref(xs, text)
    assert y_f.shape == y_u.shape
    # the query order comes from a top-k over near-tied scores: compare as sets of (box, scores) rows
    a, b = y_f.float(), y_u.float()
    dist = torch.cdist(a, b).min(-1).values
    assert (dist < 3e-2).float().mean() > 0.9


@pytest.mark.parametrize("train", [True, False])
def test_fused_glue_matches_torch_path(cuda_lib, train, monkeypatch):
    """the fold as one autograd node over the glue kernels (csrc/foldglue.cu) == the same fold as differentiable torch ops:
    values, selected rows, BatchNorm side effects and every parameter gradient"""
    import copy
    from tamtr_b200 import fold
    m, xs, text = _head()
    if not train:
        m.eval()
    ref = copy.deepcopy(m)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        tok, shapes, _ = m._encode(xs)
        assert tok.a_ext_t is not None                     # fused path taken
        monkeypatch.setattr(fold, "FUSED_GLUE", False)
        tok_r, _, _ = ref._encode(xs)
        assert tok_r.a_ext_t is None and tok_r.A is not None
        monkeypatch.setattr(fold, "FUSED_GLUE", True)
        for a, b in zip(tok.values, tok_r.values):
            assert rel_l2(a, b) < 4e-3
        assert rel_l2(tok.scores, tok_r.scores) < 2e-3
        idx = torch.randint(0, tok.B * tok.Lv, (257,), device="cuda")
        assert rel_l2(tok.rows(idx), tok_r.rows(idx)) < 2e-3
    for p, r in zip(m.input_proj, ref.input_proj):
        assert rel_l2(p[1].running_mean, r[1].running_mean) < 1e-5
        assert rel_l2(p[1].running_var, r[1].running_var) < 1e-5
        assert int(p[1].num_batches_tracked) == int(r[1].num_batches_tracked)
    # gradients through a probe on the projected values and the selected rows themselves (a whole-head loss would also
    # compare which near-tied tokens the two runs happened to select)
    def grads(model, tokens):
        for p in model.parameters():
            p.grad = None
        gen = torch.Generator(device="cuda").manual_seed(21)
        loss = sum((v.float() * torch.randn(v.shape, generator=gen, device="cuda")).sum() for v in tokens.values)
        r = tokens.rows(idx)
        loss = loss + (r.float() * torch.randn(r.shape, generator=gen, device="cuda")).sum() * 50.0
        loss.backward()
        return {n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None}
    g_f = grads(m, tok)
    monkeypatch.setattr(fold, "FUSED_GLUE", False)
    g_u = grads(ref, tok_r)
    assert set(g_f) == set(g_u) and any(k.startswith("input_proj.0.0") for k in g_f)
    for k in sorted(g_f):
        assert rel_l2(g_f[k], g_u[k]) < 5e-3, (k, rel_l2(g_f[k], g_u[k]))


def test_rank_constants_kernel_matches_torch(cuda_lib):
    """tamtr_fold_rank_consts == the torch construction of the ranking operand and constants (fp32 and bf16 parameters)"""
    from tamtr_b200 import fold
    torch.manual_seed(4)
    d, nc = 512, 10
    lin = torch.nn.Linear(d, d).cuda()
    ln = torch.nn.LayerNorm(d).cuda()
    torch.nn.init.uniform_(ln.weight, 0.5, 1.5)
    torch.nn.init.uniform_(ln.bias, -0.5, 0.5)
    sc = torch.nn.Linear(d, nc).cuda()
    for lowp in (False, True):
        a, b = (lin.bfloat16(), sc.bfloat16()) if lowp else (lin, sc)
        ref = fold._rank_constants(a, ln, b, torch.float32, fused=True)
        got = fold._rank_constants_fused(a, ln, b, True)
        assert got["npad"] == ref["npad"] == 16 and got["fused"]
        we_all = torch.cat([a.weight.float(), ref["Wr"]], 0)
        assert rel_l2(got["We_all"], we_all) < 2e-3          # the tail rows come out of a TF32 product
        assert torch.allclose(got["consts"], ref["consts"], rtol=1e-4, atol=1e-4), (got["consts"] - ref["consts"]).abs().max()


def test_feature_maps_that_require_a_gradient(cuda_lib):
    """real training: the backbone's maps require a gradient -> the folded path (differentiable torch glue) hands
    d(loss)/d(maps) back through the projection, the statistics and the selected rows (fused glue: assembled in fp32 by
    fold._maps_gradient; torch glue: by autograd).  Judged like the parameter
    gradients: against an fp32 run of the unfolded kernels, the folded bf16 result must not be further away than the
    unfolded bf16 one (with the seeded weights both sit ~20 % from fp32: d(maps) is what is left after BatchNorm's
    backward has projected the batch mean and variance directions out)."""
    import copy
    m, xs, text = _head()
    ref = copy.deepcopy(m)
    ref.folded_projection = False
    ref32 = copy.deepcopy(ref).float()

    def run(model, maps, autocast=True):
        maps = [x.detach().clone().requires_grad_() for x in maps]
        for p in model.parameters():
            p.grad = None
        torch.manual_seed(3)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            tok, shapes, hub = model._encode(maps)
            out = model(maps, text)
        db, ds, eb, es = out[:4]
        loss = db.float().square().mean() + 0.1 * ds.float().sigmoid().mean() + eb.float().square().mean() + 0.1 * es.float().sigmoid().mean()
        loss.backward()
        return tok, [x.grad.float() for x in maps]
    from tamtr_b200 import fold
    tok, gx_f = run(m, xs)
    assert getattr(tok, "is_folded", False) and tok.a_ext_t is not None    # folded, fused glue
    try:
        fold.FUSED_GLUE = False
        tok_g, gx_g = run(m, xs)                                            # same module, torch glue
        assert tok_g.a_ext_t is None and tok_g.A is not None
    finally:
        fold.FUSED_GLUE = True
    _, gx_u = run(ref, xs)
    _, gx_32 = run(ref32, [x.float() for x in xs], autocast=False)
    for a, g, b, c in zip(gx_f, gx_g, gx_u, gx_32):
        e_f, e_g, e_u = rel_l2(a, c), rel_l2(g, c), rel_l2(b, c)
        assert a.shape == c.shape and e_f < max(1.25 * e_u, 3e-2), (e_f, e_u)
        assert e_g < max(1.25 * e_u, 3e-2), (e_g, e_u)


@pytest.mark.parametrize("nbytes,n_ctas", [(16, 4), (32768, 1), (32768 * 5 + 48, 3), (32768 * 257 + 7, 32), (1 << 20, 148),
                                           (7, -2), (16, -1), (32768 * 5 + 48, -3), (32768 * 257 + 7, -148)])
def test_zero_fill_background(cuda_lib, nbytes, n_ctas):
    """tamtr_zero_fill_background zeroes exactly [ptr, ptr + bytes): whole 32 KB tiles by bulk stores, the ragged tail by
    byte stores, nothing before or after (n_ctas < 0: the variant storing 16 bytes at a time from registers)."""
    from tamtr_b200 import _lib
    pad = 256
    buf = torch.full((pad + nbytes + pad,), 0xAB, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().tamtr_zero_fill_background(buf.data_ptr() + pad, nbytes, n_ctas, _lib.stream_ptr(buf.device)),
               "zero_fill_background")
    torch.cuda.synchronize()
    assert int(buf[pad:pad + nbytes].max()) == 0
    assert int(buf[:pad].min()) == 0xAB and int(buf[pad + nbytes:].min()) == 0xAB


def test_arena_zero_fill_forked_beside_the_decoder_is_the_same_step(cuda_lib, monkeypatch):
    """ValueArena.prefill (zero fill of the samplers' gradient buffer on a side stream, forked after the projection and
    joined by the first sampler backward) against the memset in the backward: same loss, and gradients equal up to the
    order of the samplers' atomic adds (the captured step, where fork and join become branches of the graph, is covered by
    tests/test_step_gpu.py: the forked fill is the default)."""
    import copy
    from tamtr_b200 import ops
    m, xs, text = _head()
    ref = copy.deepcopy(m)
    monkeypatch.setattr(ops, "ARENA_PREFILL", False)
    loss_0, g_0 = _step(ref, xs, text)
    monkeypatch.setattr(ops, "ARENA_PREFILL", True)
    for ctas, kernel in ((0, "bulk"), (32, "bulk"), (0, "regs"), ("memset", "bulk")):
        monkeypatch.setattr(ops, "ARENA_FILL_CTAS", ctas)
        monkeypatch.setattr(ops, "ARENA_FILL_KERNEL", kernel)
        mm = copy.deepcopy(m)
        loss_1, g_1 = _step(mm, xs, text)
        assert loss_1.item() == loss_0.item()
        assert set(g_1) == set(g_0)
        for k in g_0:
            assert rel_l2(g_1[k], g_0[k]) < 2e-2, k
