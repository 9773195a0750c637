"""Algebra of the folded encoder side (tamtr_b200/fold.py) against the unfolded modules, in float64 on the CPU.

The two CUDA kernels behind fold.py are replaced here by their one-line definitions (the hooks fold._kernel_reduce /
fold._kernel_project), so what is checked is everything around them: BatchNorm statistics from the moments of X, the folded
weights and biases, the running-statistics update, the row recomputation, and every gradient autograd carries through the
fold (conv weight, gamma, beta, value_proj, the feature maps).  The kernels themselves are checked on the GPU
(tests/test_fold_gpu.py)."""
import types

import pytest
import torch
import torch.nn as nn

from tamtr_b200 import _lib, fold


def _reduce(a, a_row, a_img, token_major, x, M):
    B, C = x.shape[:2]
    X = x.flatten(2)                                          # [B, C, HW]
    A = a[:, :X.shape[2]].transpose(1, 2) if token_major else a.flatten(2)    # [B, M, HW]
    return torch.einsum("bmt,bct->mc", A, X), A.sum((0, 2))


def _project(x, w, bias, out0, out1, raw, start, N0, N1, NT, rank=None, zero=None):
    X = x.flatten(2).transpose(1, 2)                          # [B, HW, C]
    y = X @ w.t() + bias
    hw = X.shape[1]
    out0[:, start:start + hw] = y[..., :N0]
    if zero is not None:
        zero[:, start:start + hw] = 0
    if rank is None:
        out1[:, start:start + hw] = y[..., N0:N0 + N1]
        raw[:, start:start + hw] = y[..., N0 + N1:]
        return
    # the epilogue of tamtr_tok_project_rank (include/tamtr_b200.h), stated with the same sums
    scores, valid_u8, consts, nc, eps = rank
    ok = valid_u8[start:start + hw].bool().view(1, hw, 1)
    E, tail = y[..., N0:N0 + N1] * ok, y[..., N0 + N1:] * ok
    d = N1
    mean = (E.sum(-1) + consts[0]) / d
    var = ((E * E).sum(-1) + 2 * tail[..., NT - 1] + consts[1]) / d - mean * mean
    rstd = torch.rsqrt(var.clamp_min(0) + eps)
    bw, sw, ck = consts[2:2 + NT], consts[2 + NT:2 + 2 * NT], consts[2 + 2 * NT:2 + 3 * NT]
    s = rstd.unsqueeze(-1) * (tail + bw - mean.unsqueeze(-1) * sw) + ck
    scores[:, start:start + hw] = s[..., :nc].max(-1).values


@pytest.fixture
def cpu_fold(monkeypatch):
    monkeypatch.setattr(fold, "_kernel_reduce", _reduce)
    monkeypatch.setattr(fold, "_kernel_project", _project)
    monkeypatch.setattr(fold, "MATH_DTYPE", torch.float64)
    monkeypatch.setattr(_lib, "zeros_like_fast", torch.zeros_like)


def _setup(training, seed=0):
    torch.manual_seed(seed)
    d, chans, sizes, B, nl, nc = 16, (8, 24, 16), ((6, 4), (3, 4), (2, 2)), 3, 2, 5
    projs = nn.ModuleList(nn.Sequential(nn.Conv2d(c, d, 1, bias=False), nn.BatchNorm2d(d)) for c in chans).double()
    for p in projs:
        nn.init.uniform_(p[1].weight, 0.5, 1.5)
        nn.init.uniform_(p[1].bias, -0.5, 0.5)
        p[1].running_mean.uniform_(-0.3, 0.3)
        p[1].running_var.uniform_(0.5, 2.0)
    projs.train(training)
    attns = [types.SimpleNamespace(value_proj=nn.Linear(d, d).double(), n_heads=4) for _ in range(nl)]
    enc_linear, enc_norm, score = nn.Linear(d, d).double(), nn.LayerNorm(d).double(), nn.Linear(d, nc).double()
    nn.init.uniform_(enc_norm.weight, 0.5, 1.5)
    nn.init.uniform_(enc_norm.bias, -0.5, 0.5)
    xs = [(torch.randn(B, c, h, w, dtype=torch.float64) + 0.7).requires_grad_() for c, (h, w) in zip(chans, sizes)]
    return d, B, projs, attns, enc_linear, enc_norm, score, xs


def _unfolded(projs, attns, xs):
    feats = torch.cat([p(x).flatten(2).permute(0, 2, 1) for p, x in zip(projs, xs)], 1)
    return feats, [a.value_proj(feats) for a in attns]


@pytest.mark.parametrize("training", [True, False])
def test_folded_values_rows_and_gradients(cpu_fold, training):
    import copy
    d, B, projs, attns, enc_linear, enc_norm, score, xs = _setup(training)
    ref_projs = copy.deepcopy(projs)
    ref_attns = [types.SimpleNamespace(value_proj=copy.deepcopy(a.value_proj), n_heads=4) for a in attns]
    ref_xs = [x.detach().clone().requires_grad_() for x in xs]
    feats, ref_vals = _unfolded(ref_projs, ref_attns, ref_xs)

    tok = fold.FoldedTokens(xs, projs, training)
    vals = tok.project(attns, enc_linear, enc_norm, score)
    Lv = feats.shape[1]
    for v, r in zip(vals, ref_vals):
        assert torch.allclose(v.reshape(B, Lv, d), r, atol=1e-10, rtol=1e-9)
    # ranking side outputs: E = feats @ We^T (no bias), raw = E @ (score.weight * ln.weight)^T padded to 16 columns
    E_ref = feats @ enc_linear.weight.t()
    assert torch.allclose(tok.E, E_ref, atol=1e-10, rtol=1e-9)
    Wp = score.weight * enc_norm.weight
    assert torch.allclose(tok.raw.view(B, Lv, -1)[..., :Wp.shape[0]], E_ref @ Wp.t(), atol=1e-10, rtol=1e-9)
    assert tok.raw.shape[1] == 16 and torch.all(tok.raw[:, Wp.shape[0]:].abs() < 1e-12)
    # BatchNorm side effects
    for p, r in zip(projs, ref_projs):
        assert torch.allclose(p[1].running_mean, r[1].running_mean, atol=1e-12)
        assert torch.allclose(p[1].running_var, r[1].running_var, atol=1e-12)
        assert int(p[1].num_batches_tracked) == int(r[1].num_batches_tracked)
    # selected rows
    idx = torch.tensor([0, 5, 23, 24, 35, 36, 39, Lv + 1, 2 * Lv + 38, 3 * Lv - 1])
    rows = tok.rows(idx)
    assert torch.allclose(rows, feats.reshape(-1, d)[idx], atol=1e-10, rtol=1e-9)
    # gradients through values (dense) and rows (sparse)
    g = torch.Generator().manual_seed(3)
    cv = [torch.randn(v.shape, generator=g, dtype=torch.float64) for v in vals]
    cr = torch.randn(rows.shape, generator=g, dtype=torch.float64)
    loss = sum((v * c).sum() for v, c in zip(vals, cv)) + (rows * cr).sum()
    loss.backward()
    ref_loss = sum((r * c.reshape(r.shape)).sum() for r, c in zip(ref_vals, cv)) + (feats.reshape(-1, d)[idx] * cr).sum()
    ref_loss.backward()
    assert torch.allclose(loss, ref_loss, rtol=1e-10)
    for p, r in zip(projs, ref_projs):
        for a, b in ((p[0].weight, r[0].weight), (p[1].weight, r[1].weight), (p[1].bias, r[1].bias)):
            assert torch.allclose(a.grad, b.grad, atol=1e-8, rtol=1e-7), (a.grad - b.grad).abs().max()
    for a, r in zip(attns, ref_attns):
        assert torch.allclose(a.value_proj.weight.grad, r.value_proj.weight.grad, atol=1e-8, rtol=1e-7)
        assert torch.allclose(a.value_proj.bias.grad, r.value_proj.bias.grad, atol=1e-8, rtol=1e-7)
    for x, r in zip(xs, ref_xs):
        assert torch.allclose(x.grad, r.grad, atol=1e-8, rtol=1e-7), (x.grad - r.grad).abs().max()


def test_no_graph_projection(cpu_fold):
    d, B, projs, attns, enc_linear, enc_norm, score, xs = _setup(False)
    with torch.no_grad():
        feats, ref_vals = _unfolded(projs, attns, xs)
        tok = fold.FoldedTokens([x.detach() for x in xs], projs, False)
        vals = tok.project(attns, enc_linear, enc_norm, score)
    assert tok.arena is None
    for v, r in zip(vals, ref_vals):
        assert torch.allclose(v.reshape(r.shape), r, atol=1e-10, rtol=1e-9)


def test_fused_ranking_scores(cpu_fold):
    """the scores the projection's epilogue computes from (sum E, sum E^2, E . enc_bias) are the reference's ranking
    enc_score_head(LayerNorm(enc_output.0(valid * feats))).max(-1) (head.py:1229-1237)"""
    d, B, projs, attns, enc_linear, enc_norm, score, xs = _setup(False)
    with torch.no_grad():
        feats, _ = _unfolded(projs, attns, xs)
        Lv = feats.shape[1]
        valid = (torch.rand(Lv, generator=torch.Generator().manual_seed(9)) > 0.3)
        ref = score(enc_norm(enc_linear(valid.view(1, Lv, 1) * feats))).max(-1).values
        tok = fold.FoldedTokens([x.detach() for x in xs], projs, False)
        tok.project(attns, enc_linear, enc_norm, score, valid.to(torch.uint8))
    assert tok.E is None and tok.raw is None
    assert torch.allclose(tok.scores, ref, atol=1e-9, rtol=1e-8), (tok.scores - ref).abs().max()
