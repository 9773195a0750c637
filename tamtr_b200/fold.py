"""Folded encoder side of the detection heads: input_proj + BatchNorm + the first Linear of every consumer as ONE
projection per pyramid level, straight from the backbone's NCHW maps.

Reference chain (ultralytics/nn/modules/head.py:1202-1264, transformer.py:273, 870), per level l with X_l [B, C_l, HW_l]:

    Y = Conv2d(C_l, d, 1, bias=False)(X_l)            head.py:1206   (tokens x d)
    M = BatchNorm2d(d)(Y)                             head.py:1206   -> `feats`
    value_i = layers[i].cross_attn.value_proj(M)      transformer.py:273, once per decoder layer on the same feats
    E = enc_output[0](valid * M)                      head.py:1229, ranked by enc_score_head(LayerNorm(E)).max(-1)

BatchNorm with known statistics is affine per channel: M = Y * s + t,  s = gamma * rstd,  t = beta - mu * s.  Every
consumer starts with a Linear, so with A = diag(s) Wc  [d, C_l]:

    value_i = X_l^T (W_i A)^T + (W_i t + b_i)         K = C_l (128 / 256 / 512) instead of d = 512, no BatchNorm pass
    E       = X_l^T (W_e A)^T +  W_e t  (+ b_e)

and the batch statistics of Y follow from the first two moments of X_l (Y = Wc x):

    mu = Wc mean(x),     var = diag(Wc Cov(x) Wc^T),  Cov(x) = E[x x^T] - mean mean^T

So `feats` is never materialised: per level one second-moment reduction over X_l (tamtr_tok_reduce, C_l x C_l), a handful
of [d, C_l]-sized matrix products (differentiable torch ops: autograd carries the gradients of conv weight, gamma, beta,
value_proj and -- through the selected rows -- enc_output), one projection kernel (tamtr_tok_project) writing the value
tensors of all decoder layers, the ranking embedding and the ranking scores, and in the backward ONE reduction
grad_value^T X_l (tamtr_tok_reduce) in place of value_proj's dgrad, the BatchNorm backward passes and the conv's wgrad.
At TAM-TR shapes (B = 16, 33 600 tokens, d = 512) the dense work drops from 2.8 TFLOP + 6 passes over [B, Lv, d] to
0.85 TFLOP and no pass.  The rows picked by the query selection (B * nq of B * Lv) are recomputed from X_l through the same
A and t, differentiably (head.py:1240-1254 only ever sends gradient to those rows).

Numerics: X_l and the folded weights are bf16 operands of fp32-accumulating tensor-core products; statistics and the
fold itself are fp32.  Compared with the reference under autocast (Y and M each rounded to bf16) the folded path rounds
once less.  It is taken for bf16 activations only; fp32 inputs keep the unfolded kernels (ops.input_proj_tokens).
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import _lib, ops

MATH_DTYPE = torch.float32          # dtype of the fold (tests run the algebra in float64 on the CPU through the hooks below)


# ---------------------------------------------------------------------------------------- kernel hooks (C ABI, CUDA only)
def _kernel_reduce_parts(a, a_row, a_img, token_major, x, M):
    """tamtr_tok_reduce: per-split partials (part_d [S, M, C], part_rs [S, M]) of
    D [M, C] = sum over (image, token) of a[m; b, tok] * x[b, c, tok] and rs [M] = row sums of a."""
    _lib.require_cuda(a, x)
    if a.dtype != torch.bfloat16 or x.dtype != torch.bfloat16:
        raise RuntimeError("tamtr_b200: tok_reduce takes bf16 operands")
    B, C = x.shape[0], x.shape[1]
    HW = x.shape[2] * x.shape[3]
    lib = _lib.lib()
    S = lib.tamtr_tok_reduce_splits(B, C, HW, M, int(token_major))
    if S <= 0:
        raise RuntimeError(f"tamtr_b200: tok_reduce does not support B={B} C={C} HW={HW} M={M}")
    part_d = torch.empty(S, M, C, dtype=torch.float32, device=x.device)
    part_rs = torch.empty(S, M, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.tamtr_tok_reduce(a.data_ptr(), a_row, a_img, int(token_major), x.data_ptr(), part_d.data_ptr(),
                                  part_rs.data_ptr(), B, C, HW, M, _lib.stream_ptr(x.device))
    _lib.check(rc, "tok_reduce")
    return part_d, part_rs


def _kernel_reduce(a, a_row, a_img, token_major, x, M):
    part_d, part_rs = _kernel_reduce_parts(a, a_row, a_img, token_major, x, M)
    return part_d.sum(0), part_rs.sum(0)


def _kernel_project(x, w, bias, out0, out1, raw, start, N0, N1, NT, rank=None, zero=None):
    """out0[:, start:start+HW] = x^T w[:N0]^T + bias[:N0] (bf16).  The N1 + NT remaining columns are the ranking branch:
    stored (out1 bf16, raw fp32) when `rank` is None, else reduced in the kernel's epilogue to one score per token,
    rank = (scores [B, Lv] fp32, valid_u8 [Lv], consts, nc, eps)  (tamtr_tok_project / tamtr_tok_project_rank).
    `zero`: a second tensor like out0 whose same rows are filled with zeros (the samplers' gradient arena)."""
    _lib.require_cuda(x, w, bias, out0)
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16 or bias.dtype != torch.float32:
        raise RuntimeError("tamtr_b200: tok_project takes bf16 operands and an fp32 bias")
    B, C = x.shape[0], x.shape[1]
    HW = x.shape[2] * x.shape[3]
    Lv = out0.shape[1]
    es0 = out0.element_size()
    if zero is not None and (zero.shape != out0.shape or zero.dtype != out0.dtype or not zero.is_contiguous()):
        raise RuntimeError("tamtr_b200: the zero-filled tensor must be laid out like the projected values")
    zp = None if zero is None else zero.data_ptr() + start * N0 * es0
    with torch.cuda.device(x.device):
        if rank is None:
            rc = _lib.lib().tamtr_tok_project(
                x.data_ptr(), w.data_ptr(), bias.data_ptr(),
                out0.data_ptr() + start * N0 * es0, zp, N0, Lv * N0,
                (out1.data_ptr() + start * N1 * out1.element_size()) if N1 else None, N1, Lv * N1,
                (raw.data_ptr() + start * NT * 4) if NT else None, NT, Lv * NT,
                B, C, HW, N0, N1, NT, _lib.stream_ptr(x.device))
        else:
            scores, valid_u8, consts, nc, eps = rank
            rc = _lib.lib().tamtr_tok_project_rank(
                x.data_ptr(), w.data_ptr(), bias.data_ptr(), out0.data_ptr() + start * N0 * es0, zp, N0, Lv * N0,
                scores.data_ptr() + start * 4, Lv, valid_u8.data_ptr() + start, consts.data_ptr(), int(nc), float(eps),
                B, C, HW, N0, N1, NT, _lib.stream_ptr(x.device))
    _lib.check(rc, "tok_project")


def supported(xs, d, n_value_cols, n_tail=16):
    """True when every level can go through the folded kernels (else the caller keeps the unfolded path)."""
    if not xs or not all(x.is_cuda and x.dim() == 4 for x in xs):
        return False
    lib = _lib.lib()
    B = xs[0].shape[0]
    for x in xs:
        C, HW = x.shape[1], x.shape[2] * x.shape[3]
        if x.shape[0] != B or not lib.tamtr_tok_project_supported(B, C, HW, n_value_cols, d, n_tail):
            return False
        if not lib.tamtr_tok_reduce_supported(B, C, HW, C, 0) or not lib.tamtr_tok_reduce_supported(B, C, HW, n_value_cols, 1):
            return False
    return True


# ---------------------------------------------------------------------------------------- moments of X_l
class _TokStatsFn(torch.autograd.Function):
    """x [B, C, H, W] bf16 -> (S1 [C] = sum over tokens of x, G [C, C] = sum over tokens of x x^T), fp32."""

    @staticmethod
    def forward(ctx, x):
        B, C, H, W = x.shape
        G, S1 = _kernel_reduce(x, H * W, C * H * W, False, x, C)
        ctx.save_for_backward(x)
        return S1, G

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dS1, dG):
        (x,) = ctx.saved_tensors            # only reached when the feature maps themselves require a gradient
        B, C, H, W = x.shape
        with torch.autocast("cuda", enabled=False):
            sym = (dG + dG.t()).to(x.dtype)
            dx = torch.matmul(sym, x.flatten(2)) + dS1.to(x.dtype).view(1, C, 1)
        return dx.view(B, C, H, W)


def _moments(x):
    if x.requires_grad and torch.is_grad_enabled():
        return _TokStatsFn.apply(x)
    B, C, H, W = x.shape
    G, S1 = _kernel_reduce(x, H * W, C * H * W, False, x, C)
    return S1, G


class _FoldMatmulFn(torch.autograd.Function):
    """W [N, d] @ A [L, d, K] -> [L, N, K] for the folded weights, forward and backward on the TF32 tensor-core path when
    the operands are fp32 CUDA tensors: the products are rounded to bf16 right after (they are the projection kernel's
    weight operand), and their gradients feed parameters whose reference gradients come out of bf16 GEMMs, so 10 mantissa
    bits in the multiplications lose nothing -- while the fp32 SIMT GEMMs cost 0.4 ms per step at TAM-TR shapes."""

    @staticmethod
    def _mm(a, b):
        if not (a.is_cuda and a.dtype == torch.float32):
            return torch.matmul(a, b)
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            return torch.matmul(a, b)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old

    @staticmethod
    def forward(ctx, W, A):
        ctx.save_for_backward(W, A)
        return _FoldMatmulFn._mm(W, A)

    @staticmethod
    def backward(ctx, g):
        W, A = ctx.saved_tensors
        dW = dA = None
        if ctx.needs_input_grad[0]:
            L, N, K = g.shape
            # sum over levels of g_l @ A_l^T as one product over the concatenated (level, column) axis
            dW = _FoldMatmulFn._mm(g.permute(1, 0, 2).reshape(N, L * K), A.permute(0, 2, 1).reshape(L * K, -1))
        if ctx.needs_input_grad[1]:
            dA = _FoldMatmulFn._mm(W.t(), g)
        return dW, dA


def _fold_matmul(W, A):
    if torch.is_grad_enabled() and (W.requires_grad or A.requires_grad):
        return _FoldMatmulFn.apply(W, A)
    return _FoldMatmulFn._mm(W, A)


def _pad_cols(t, width):
    return t if t.shape[-1] == width else F.pad(t, (0, width - t.shape[-1]))


# ---------------------------------------------------------------------------------------- the projection
class _TokProjectFn(torch.autograd.Function):
    """values of all decoder layers (+ ranking embedding and scores as side outputs on `tokens`) from the folded weights.

    Fv [L, N0, Cm + 1]: per level the folded value weights (columns [0, C_l)) and the folded bias part W t (column Cm);
    bv [N0] the layers' own biases; Fe [L, N1 + NT, Cm + 1] likewise for the ranking branch (no gradient: head.py's dense
    ranking pass only selects rows).  Backward: one reduction per level over the shared gradient arena."""

    @staticmethod
    def forward(ctx, tokens, arena, n_layers, n_heads, Fv, bv, Fe, *xs):
        L, N0, Cm1 = Fv.shape
        Cm = Cm1 - 1
        N1, NT = tokens.d, Fe.shape[1] - tokens.d
        W = torch.cat([Fv, Fe], 1)                                             # [L, N_all, Cm + 1]
        bias = W[:, :, Cm].clone()
        bias[:, :N0] += bv.to(bias.dtype)
        lp = tokens.dtype                                                      # bf16 (the kernels take nothing else)
        acc = torch.float32 if lp == torch.bfloat16 else W.dtype
        bias = bias.to(acc).contiguous()
        Wb = W[:, :, :Cm].to(lp)
        ws = [Wb[l, :, :x.shape[1]].contiguous() for l, x in enumerate(xs)]
        values = _project_levels(tokens, xs, ws, bias, N0, N1, NT, n_layers, n_heads,
                                 None if isinstance(ctx, _NoCtx) else arena)
        value_all = tokens._value_all
        d = N0 // n_layers
        ctx.save_for_backward(*xs, *[w[:N0] for w in ws])
        ctx.arena, ctx.n, ctx.d, ctx.L, ctx.Cm = arena, n_layers, d, L, Cm
        ctx.meta = (value_all.shape, tokens.starts, tokens.hw, Fv.dtype, bv.dtype, lp)
        ctx.set_materialize_grads(False)
        arena.base = value_all
        return values

    @staticmethod
    def backward(ctx, *grads):
        L, Cm, n, d = ctx.L, ctx.Cm, ctx.n, ctx.d
        xs, wl = ctx.saved_tensors[:L], ctx.saved_tensors[L:]
        shape, starts, hw, fdt, bdt, lp = ctx.meta
        dev = xs[0].device
        buf = _arena_gradient(ctx.arena, grads, shape, d, lp, dev)
        B, Lv, N0 = shape
        dFv = torch.zeros(L, N0, Cm + 1, dtype=fdt, device=dev)
        dxs = []
        for l, x in enumerate(xs):
            C = x.shape[1]
            a = buf[:, starts[l]:]
            D, rs = _kernel_reduce(a, N0, Lv * N0, True, x, N0)
            dFv[l, :, :C] = D
            dFv[l, :, Cm] = rs
            if ctx.needs_input_grad[7 + l]:
                with torch.autocast("cuda", enabled=False):
                    gx = torch.matmul(wl[l].t(), buf[:, starts[l]:starts[l] + hw[l]].transpose(1, 2))    # [B, C, HW]
                dxs.append(gx.view(x.shape).to(x.dtype))
            else:
                dxs.append(None)
        dbv = dFv[:, :, Cm].sum(0).to(bdt) if ctx.needs_input_grad[5] else None
        return (None, None, None, None, dFv, dbv, None, *dxs)


# ---------------------------------------------------------------------------------------- fused glue (csrc/foldglue.cu)
# Gradient arena zeroed by tamtr_tok_project's own stores instead of a memset node in the backward.  Measured: the projection
# gets slower by exactly the memset's 265 us (its extra 1.65 GB of stores are HBM-write-bound either way), so it is off.
ZERO_FILL_IN_PROJECTION = os.environ.get("TAMTR_FOLD_ZERO", "0") != "0"
FUSED_GLUE = os.environ.get("TAMTR_FOLD_GLUE", "1") != "0"    # the whole fold as ONE autograd node over a handful of
                                                             # kernels (False: the differentiable torch ops)

_MAXL = 8


def _kpad(Cm):
    """columns of the A_ext layouts of csrc/foldglue.cu: Cm weights + the bias part, rounded up to a multiple of 8"""
    return (Cm + 1 + 7) // 8 * 8


def _parr(ts):
    return (ctypes.c_void_p * _MAXL)(*([None if t is None else t.data_ptr() for t in ts] + [None] * (_MAXL - len(ts))))


def _iarr(v):
    return (ctypes.c_int * _MAXL)(*(list(map(int, v)) + [0] * (_MAXL - len(v))))


def _farr(v):
    return (ctypes.c_float * _MAXL)(*(list(map(float, v)) + [0.0] * (_MAXL - len(v))))


def _mm_tf32(a, b):
    return _FoldMatmulFn._mm(a, b)


def _mm_fp32(a, b):
    """true-fp32 product whatever the global TF32 switch says (variance of the conv output: a difference of large terms)"""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return torch.mm(a, b)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def fused_glue_applies(tokens, projs, attns):
    if not (FUSED_GLUE and tokens.is_cuda and MATH_DTYPE == torch.float32 and len(tokens.xs) <= _MAXL):
        return False
    for conv, bn in projs:
        ts = [conv.weight, bn.weight, bn.bias] + ([bn.running_mean, bn.running_var] if bn.running_mean is not None else [])
        if any(t.dtype != torch.float32 or not t.is_contiguous() for t in ts):
            return False
    stats = [bn.running_mean is None for _, bn in projs]
    return all(stats) or not any(stats)


class _FusedFoldFn(torch.autograd.Function):
    """The folded encoder side as one autograd node: moments -> BatchNorm coefficients -> folded weights -> projection
    (forward), weight-gradient reductions -> d(folded weights) -> d(value_proj), d(A) -> d(conv weight, gamma, beta)
    (backward), over tamtr_tok_reduce / tamtr_tok_project* and the glue kernels of csrc/foldglue.cu.  The products that
    stay library GEMMs: conv_w @ Cov (fp32) and value_proj / ranking weights @ A_ext (TF32), 2 + L launches.

    Outputs: a_ext_t [L, K, d] (differentiable: the selected rows are Xcat @ a_ext_t, see FoldedTokens.rows) and the
    value views of the decoder layers."""

    @staticmethod
    def forward(ctx, tokens, arena, training, bns, n_layers, n_heads, We_all, *t):
        L = len(tokens.xs)
        xs, convs, gammas, betas = t[:L], t[L:2 * L], t[2 * L:3 * L], t[3 * L:4 * L]
        wvs, bvs = t[4 * L:4 * L + n_layers], t[4 * L + n_layers:]
        cs, d, Cm = tokens.cs, tokens.d, tokens.Cm
        K, dev, lib = _kpad(Cm), tokens.device, _lib.lib()
        st = _lib.stream_ptr(dev)
        n_tok = [float(tokens.B * n) for n in tokens.hw]
        batch_stats = training or bns[0].running_mean is None
        update = batch_stats and training and bns[0].running_mean is not None
        C_arr = _iarr(cs)
        P, mean_x = [None] * L, [None] * L
        with torch.cuda.device(dev):
            if batch_stats:
                parts = [_kernel_reduce_parts(x, x.shape[2] * x.shape[3], x.shape[1] * x.shape[2] * x.shape[3], False, x,
                                              x.shape[1]) for x in xs]
                mean_flat = torch.empty(sum(cs), dtype=torch.float32, device=dev)
                cov_flat = torch.empty(sum(c * c for c in cs), dtype=torch.float32, device=dev)
                o1 = o2 = 0
                covs = []
                for l, c in enumerate(cs):
                    mean_x[l] = mean_flat[o1:o1 + c]
                    covs.append(cov_flat[o2:o2 + c * c].view(c, c))
                    o1, o2 = o1 + c, o2 + c * c
                _lib.check(lib.tamtr_fold_stats(L, C_arr, _iarr([p[0].shape[0] for p in parts]), _farr(n_tok),
                                                _parr([p[0] for p in parts]), _parr([p[1] for p in parts]),
                                                _parr(mean_x), _parr(covs), st), "fold_stats")
                P = [_mm_fp32(convs[l].view(d, cs[l]), covs[l]) for l in range(L)]
            mom = [0.0] * L
            if update:
                mom = [b.momentum if b.momentum is not None else 1.0 / float(b.num_batches_tracked + 1) for b in bns]
            a_ext = torch.empty(L, d, K, dtype=torch.float32, device=dev)
            a_ext_t = torch.empty(L, K, d, dtype=torch.float32, device=dev)
            stats = torch.empty(L, d, 4, dtype=torch.float32, device=dev)
            has_run = bns[0].running_mean is not None
            _lib.check(lib.tamtr_fold_bn(
                L, d, C_arr, _farr(n_tok), _parr([c.view(d, -1) for c in convs]), _parr(P) if batch_stats else None,
                _parr(mean_x) if batch_stats else None, _parr(gammas), _parr(betas),
                _parr([b.running_mean for b in bns]) if has_run else None,
                _parr([b.running_var for b in bns]) if has_run else None,
                _parr([b.num_batches_tracked for b in bns]) if has_run else None, _farr(mom), _farr([b.eps for b in bns]),
                int(batch_stats), int(update), a_ext.data_ptr(), a_ext_t.data_ptr(), stats.data_ptr(), st), "fold_bn")
            Wv = torch.cat(wvs, 0).float()
            bv = torch.cat(bvs, 0).float()
            Fv = _mm_tf32(Wv, a_ext)                                      # [L, N0, K]
            Fe = _mm_tf32(We_all, a_ext)                                  # [L, NE, K]
            N0, NE = Wv.shape[0], We_all.shape[0]
            w_out = [torch.empty(N0 + NE, c, dtype=torch.bfloat16, device=dev) for c in cs]
            bias = torch.empty(L, N0 + NE, dtype=torch.float32, device=dev)
            _lib.check(lib.tamtr_fold_pack(L, C_arr, Fv.data_ptr(), Fe.data_ptr(), bv.data_ptr(), _parr(w_out),
                                           bias.data_ptr(), N0, NE, st), "fold_pack")
        values = _project_levels(tokens, xs, w_out, bias, N0, tokens.d, NE - tokens.d, n_layers, n_heads,
                                 None if isinstance(ctx, _NoCtx) else arena)
        value_all = tokens._value_all
        ctx.save_for_backward(*xs, *convs, *gammas, *[p for p in P if p is not None],
                              *([mean_flat] if batch_stats else []), stats, a_ext_t, Wv, *w_out)
        ctx.arena, ctx.L, ctx.n, ctx.batch_stats = arena, L, n_layers, batch_stats
        # (no reference to `tokens` itself: it holds a_ext_t, whose grad_fn is this node -- a cycle would keep the graph alive)
        ctx.n_tok, ctx.row_grads, ctx.geom = n_tok, tokens._row_grads, (tokens.Lv, tokens.starts, tokens.hw, tokens.Cm)
        ctx.meta = (value_all.shape, tokens.starts, tokens.hw, cs, d, Cm, [w.dtype for w in wvs], [b.dtype for b in bvs],
                    [c.shape for c in convs])
        ctx.set_materialize_grads(False)
        arena.base = value_all
        return (a_ext_t, *values)

    @staticmethod
    def backward(ctx, d_aext_t, *grads):
        L, n = ctx.L, ctx.n
        shape, starts, hw, cs, d, Cm, wdts, bdts, cshapes = ctx.meta
        sv = ctx.saved_tensors
        xs, convs, gammas = sv[:L], sv[L:2 * L], sv[2 * L:3 * L]
        k = 3 * L
        P = mean_flat = None
        if ctx.batch_stats:
            P, mean_flat = sv[k:k + L], sv[k + L]
            k += L + 1
        stats, a_ext_t, Wv = sv[k], sv[k + 1], sv[k + 2]
        w_out = sv[k + 3:k + 3 + L]
        need_dx = [bool(f) for f in ctx.needs_input_grad[7:7 + L]]
        K, dev, lib = _kpad(Cm), xs[0].device, _lib.lib()
        st = _lib.stream_ptr(dev)
        buf = _arena_gradient(ctx.arena, grads, shape, Wv.shape[0] // n, torch.bfloat16, dev)
        B, Lv, N0 = shape
        C_arr = _iarr(cs)
        with torch.cuda.device(dev):
            parts = [_kernel_reduce_parts(buf[:, starts[l]:], N0, Lv * N0, True, xs[l], N0) for l in range(L)]
            dF = torch.empty(L, N0, K, dtype=torch.float32, device=dev)
            dF_t = torch.empty(N0, L, K, dtype=torch.float32, device=dev)
            d_bv = torch.empty(N0, dtype=torch.float32, device=dev)
            _lib.check(lib.tamtr_fold_unpack(L, C_arr, _iarr([p[0].shape[0] for p in parts]), _parr([p[0] for p in parts]),
                                             _parr([p[1] for p in parts]), dF.data_ptr(), dF_t.data_ptr(), d_bv.data_ptr(),
                                             N0, st), "fold_unpack")
            dWv = _mm_tf32(dF_t.view(N0, L * K), a_ext_t.view(L * K, d))
            dA = _mm_tf32(Wv.t(), dF)                                      # [L, d, K]
            d_wc = [torch.empty(d, c, dtype=torch.float32, device=dev) for c in cs]
            d_gamma = [torch.empty(d, dtype=torch.float32, device=dev) for _ in cs]
            d_beta = [torch.empty(d, dtype=torch.float32, device=dev) for _ in cs]
            mean_x = None
            if ctx.batch_stats:
                o, mean_x = 0, []
                for c in cs:
                    mean_x.append(mean_flat[o:o + c])
                    o += c
            dAt = None if d_aext_t is None else d_aext_t.contiguous()
            d_stat = torch.empty(L, d, 2, dtype=torch.float32, device=dev) if (any(need_dx) and ctx.batch_stats) else None
            _lib.check(lib.tamtr_fold_bn_bwd(
                L, d, C_arr, _parr([c.view(d, -1) for c in convs]), _parr(P) if P is not None else None,
                _parr(mean_x) if mean_x is not None else None, _parr(gammas), dA.data_ptr(),
                None if dAt is None else dAt.data_ptr(), stats.data_ptr(), int(ctx.batch_stats), _parr(d_wc),
                _parr(d_gamma), _parr(d_beta), None if d_stat is None else d_stat.data_ptr(), st), "fold_bn_bwd")
            dxs = [None] * L
            for l in range(L):
                if need_dx[l]:
                    dxs[l] = _maps_gradient(ctx.geom, ctx.row_grads, l, xs[l], w_out[l][:N0],
                                            buf[:, starts[l]:starts[l] + hw[l]], convs[l].view(d, -1),
                                            None if d_stat is None else d_stat[l], None if mean_x is None else mean_x[l],
                                            ctx.n_tok[l])
            del ctx.row_grads[:]
        dd = Wv.shape[0] // n
        dWv_c = dWv if all(t == torch.float32 for t in wdts) else dWv.to(wdts[0])
        d_bv_c = d_bv if all(t == torch.float32 for t in bdts) else d_bv.to(bdts[0])
        g_wv = [dWv_c[i * dd:(i + 1) * dd].to(wdts[i]) for i in range(n)]
        g_bv = [d_bv_c[i * dd:(i + 1) * dd].to(bdts[i]) for i in range(n)]
        return (None,) * 7 + tuple(dxs) + tuple(g.view(s) for g, s in zip(d_wc, cshapes)) + tuple(d_gamma) \
            + tuple(d_beta) + tuple(g_wv) + tuple(g_bv)


def _bmm_f32(a, b):
    """batched bf16 product with an fp32 result"""
    try:
        return torch.bmm(a, b, out_dtype=torch.float32)
    except (TypeError, RuntimeError):
        return torch.bmm(a.float(), b.float())


def _maps_gradient(geom, row_grads, l, x, w_v, g_l, wc, d_stat, mean_x, n_tok):
    """d(loss)/d(X_l) of the fused fold, assembled in fp32 and rounded once:
      W_fold^T grad_value^T                          the projection (X enters it linearly)
      (2 / n) dCov X + dS1                           through the batch statistics: mu = Wc mean(x), var = diag(Wc Cov Wc^T),
                                                     Cov = G / n - m m^T   (dCov = Wc^T diag(dvar) Wc, dmean = Wc^T dmu)
      row-sparse terms of the selected tokens        (FoldedTokens.rows: Xcat @ A_ext^T)"""
    B, C = x.shape[0], x.shape[1]
    X = x.flatten(2)                                                                  # [B, C, HW]
    gx = _bmm_f32(w_v.t().unsqueeze(0).expand(B, -1, -1), g_l.transpose(1, 2))        # [B, C, HW] fp32
    if d_stat is not None:
        dmu, dvar = d_stat[:, 0], d_stat[:, 1]
        dcov = _mm_fp32((wc * dvar.unsqueeze(1)).t().contiguous(), wc)                # [C, C]
        ds1 = (dmu @ wc - 2.0 * (dcov @ mean_x)) / n_tok
        gx = gx + _bmm_f32((dcov * (2.0 / n_tok)).to(x.dtype).unsqueeze(0).expand(B, -1, -1), X) + ds1.view(1, C, 1)
    Lv, starts, hw, Cm = geom
    K = _kpad(Cm)
    for flat_idx, dxcat in row_grads:                                                 # [R], [R, L * K] fp32
        img = torch.div(flat_idx, Lv, rounding_mode="floor")
        rel = flat_idx - img * Lv - starts[l]
        inside = ((rel >= 0) & (rel < hw[l])).to(dxcat.dtype).unsqueeze(1)
        vals = dxcat[:, l * K:l * K + C] * inside
        gx.permute(0, 2, 1).index_put_((img, rel.clamp(0, hw[l] - 1)), vals, accumulate=True)
    return gx.to(x.dtype).view(x.shape)


class _RowsFn(torch.autograd.Function):
    """feats rows = Xcat [R, L*K] @ A_ext^T [L*K, d] (TF32 forward and backward).  Xcat is gathered by a kernel; when the
    feature maps require a gradient its row-sparse gradient is left on `tokens` for _FusedFoldFn.backward, which runs
    after this node (it consumes d(A_ext^T)) and adds it into the dense d(maps)."""

    @staticmethod
    def forward(ctx, tokens, flat_idx, xcat, a_ext_t):
        ctx.save_for_backward(xcat, a_ext_t, flat_idx)
        ctx.row_grads = tokens._row_grads if tokens.maps_need_grad else None     # (the list, not `tokens`: no cycle)
        return _mm_tf32(xcat, a_ext_t.reshape(-1, a_ext_t.shape[-1]))

    @staticmethod
    def backward(ctx, g):
        xcat, a_ext_t, flat_idx = ctx.saved_tensors
        g = g.float().contiguous()
        if ctx.row_grads is not None:
            ctx.row_grads.append((flat_idx, _mm_tf32(g, a_ext_t.reshape(-1, a_ext_t.shape[-1]).t())))
        return None, None, None, _mm_tf32(xcat.t(), g).view(a_ext_t.shape)


def _arena_gradient(arena, grads, shape, d, lp, dev):
    """The gradient of the projected values: what the samplers accumulated in the shared arena (+ whatever reached a
    value view through plain autograd), as one [B, Lv, N0] tensor of dtype `lp`."""
    buf, written = arena.take()
    arena.bias_grad = {}
    if buf is None:
        buf = _lib.zeros_like_fast(torch.empty(shape, dtype=lp, device=dev))
    for i, g in enumerate(grads):
        if g is None:
            continue
        if i * d in written:
            if any(st != 0 for st in g.stride()):
                raise RuntimeError("tamtr_b200: a projected value view has a consumer besides its sampler")
            continue
        buf[:, :, i * d:(i + 1) * d].add_(g.reshape(shape[0], shape[1], d))
    if buf.dtype != lp:                             # fp32 arena: one cast pass in front of the reductions
        buf = buf.to(lp)
    return buf


def _project_levels(tokens, xs, ws, bias, N0, N1, NT, n_layers, n_heads, arena=None):
    """One tamtr_tok_project(_rank) launch per level into the value tensor of all layers (+ the ranking side outputs on
    `tokens`); returns the per-layer views [B, Lv, H, Dh].  `arena` (a backward will follow): the samplers' gradient buffer
    is allocated here and zero-filled by the same kernels -- its 1.65 GB memset node per step disappears."""
    B, Lv, dev, lp = tokens.B, tokens.Lv, xs[0].device, tokens.dtype
    acc = torch.float32 if lp == torch.bfloat16 else bias.dtype
    value_all = torch.empty(B, Lv, N0, dtype=lp, device=dev)
    zero = None
    if arena is not None and ZERO_FILL_IN_PROJECTION and lp == torch.bfloat16 and arena.grad_dtype in (None, lp) \
            and arena.buf is None:
        zero = torch.empty_like(value_all)
    rk = tokens.rank_consts
    if rk["fused"]:        # ranking finished in the projection's epilogue: E and the class scores are never stored
        E = raw = None
        scores = torch.empty(B, Lv, dtype=acc, device=dev)
        rank = (scores, tokens.valid_u8, rk["consts"], rk["nc"], rk["eps"])
    else:
        E = torch.empty(B, Lv, N1, dtype=lp, device=dev)
        raw = torch.empty(B * Lv, NT, dtype=acc, device=dev)
        scores = rank = None
    for l, x in enumerate(xs):
        _kernel_project(x, ws[l], bias[l], value_all, E, None if raw is None else raw.view(B, Lv, NT), tokens.starts[l],
                        N0, N1, NT, rank, zero)
    if zero is not None:
        arena.buf = zero
    elif arena is not None:
        arena.prefill(value_all.shape, lp, dev)     # zero fill forked here, beside the (latency-bound) decoder forward
    tokens.E, tokens.raw, tokens.scores, tokens._value_all = E, raw, scores, value_all
    d = N0 // n_layers
    return tuple(value_all[:, :, i * d:(i + 1) * d].view(B, Lv, n_heads, d // n_heads) for i in range(n_layers))


# ---------------------------------------------------------------------------------------- the token source
class FoldedTokens:
    """Stands where `feats` [B, Lv, d] stands in the heads when the encoder side is folded: carries the NCHW maps and the
    per-level affine map  M = X^T A^T + t  instead of the materialised tokens."""

    is_folded = True

    def __init__(self, xs, projs, training):
        self.xs = [x.contiguous() for x in xs]
        self.shapes = [[x.shape[2], x.shape[3]] for x in xs]
        self.hw = [h * w for h, w in self.shapes]
        self.starts = [sum(self.hw[:i]) for i in range(len(xs))]
        self.B, self.Lv = xs[0].shape[0], sum(self.hw)
        self.device, self.dtype = xs[0].device, xs[0].dtype
        self.cs = [x.shape[1] for x in xs]
        self.Cm = max(self.cs)
        self.d = projs[0][0].weight.shape[0]
        self.shape = (self.B, self.Lv, self.d)
        self.is_cuda = xs[0].is_cuda
        self.values = self.arena = self.E = self.raw = self.scores = self.valid_u8 = None
        self.projs, self.training = projs, training
        self.maps_need_grad, self._row_grads = False, []
        self.A = self.t = self.a_ext_t = None       # set by project(): A [L, d, Cm], t [L, d]  or  a_ext_t [L, Cm + 1, d]

    # BatchNorm2d's forward (torch/nn/modules/batchnorm.py:155-193) on statistics derived from the moments of X
    def _coefficients(self, projs, training):
        md, Cm, d, L = MATH_DTYPE, self.Cm, self.d, len(self.xs)
        convs, bns = [p[0] for p in projs], [p[1] for p in projs]
        Wc = torch.stack([_pad_cols(c.weight.reshape(d, -1).to(md), Cm) for c in convs])             # [L, d, Cm]
        gamma = torch.stack([b.weight.to(md) for b in bns])
        beta = torch.stack([b.bias.to(md) for b in bns])
        eps = bns[0].eps            # python scalars only: nothing here may copy from the host (CUDA-graph capture)
        if any(b.eps != eps for b in bns):
            eps = torch.stack([torch.full((1,), b.eps, dtype=md, device=self.device) for b in bns])
        batch_stats = [training or b.running_mean is None for b in bns]
        if any(batch_stats) and not all(batch_stats):
            raise RuntimeError("tamtr_b200: input_proj BatchNorm layers must agree on track_running_stats")
        if batch_stats[0]:
            # per level, unpadded and in full fp32: var = diag(Wc Cov Wc^T) is a difference of large terms
            mus, vars_ = [], []
            for l, (x, C) in enumerate(zip(self.xs, self.cs)):
                S1, G = _moments(x)
                n_tok = float(x.shape[0] * x.shape[2] * x.shape[3])
                mean_x = S1.to(md) / n_tok
                cov = torch.addcmul(G.to(md) / n_tok, mean_x.unsqueeze(1), mean_x.unsqueeze(0), value=-1.0)
                w = Wc[l, :, :C]
                mus.append(w @ mean_x)
                vars_.append(((w @ cov) * w).sum(-1))
            mu = torch.stack(mus)
            var = torch.stack(vars_).clamp_min(0)
            if training:
                with torch.no_grad():
                    for l, b in enumerate(bns):
                        if b.running_mean is None:
                            continue
                        mom = b.momentum if b.momentum is not None else 1.0 / float(b.num_batches_tracked + 1)
                        n_tok = float(self.B * self.hw[l])
                        b.running_mean.mul_(1 - mom).add_(mu[l].to(b.running_mean.dtype), alpha=mom)
                        b.running_var.mul_(1 - mom).add_((var[l] * (n_tok / max(n_tok - 1, 1))).to(b.running_var.dtype),
                                                         alpha=mom)
                        b.num_batches_tracked += 1
        else:
            mu = torch.stack([b.running_mean.to(md) for b in bns])
            var = torch.stack([b.running_var.to(md) for b in bns])
        s = gamma * torch.rsqrt(var + eps)
        t = beta - mu * s
        return s.unsqueeze(-1) * Wc, t

    # --- dense side: values of every decoder layer, ranking embedding and scores, one kernel per level
    def project(self, attns, enc_linear, enc_norm, score_linear, valid_u8=None):
        """`valid_u8` [Lv]: the anchors' validity mask (head.py:1200); with it (and at most 62 classes) the ranking scores
        come straight out of the projection kernel, otherwise rank() runs tamtr_rank_tokens on the stored E / scores."""
        md = MATH_DTYPE
        self.valid_u8 = valid_u8
        n_layers, n_heads = len(attns), attns[0].n_heads
        grad = torch.is_grad_enabled()
        with torch.autocast(self.device.type, enabled=False):
            with torch.no_grad():
                rk = _rank_constants_fused(enc_linear, enc_norm, score_linear, valid_u8 is not None) \
                    if (FUSED_GLUE and self.is_cuda and md == torch.float32) else None
                if rk is None:
                    rk = _rank_constants(enc_linear, enc_norm, score_linear, md, fused=valid_u8 is not None)
                    rk["We_all"] = torch.cat([enc_linear.weight.to(md), rk["Wr"]], 0)
                We_all = rk["We_all"]
            self.rank_consts = rk
            self.arena = ops.ValueArena()
            self.arena.defer = True         # the heads fork the zero fill after their query selection (start_prefill)
            if fused_glue_applies(self, self.projs, attns):
                convs, bns = [p[0] for p in self.projs], [p[1] for p in self.projs]
                t = list(self.xs) + [c.weight for c in convs] + [b.weight for b in bns] + [b.bias for b in bns] \
                    + [a.value_proj.weight for a in attns] + [a.value_proj.bias for a in attns]
                self.maps_need_grad = grad and any(x.requires_grad for x in self.xs)
                self._row_grads = []
                if grad and any(p.requires_grad for p in t):
                    out = _FusedFoldFn.apply(self, self.arena, self.training, bns, n_layers, n_heads, We_all, *t)
                else:
                    out = _FusedFoldFn.forward(_NoCtx(), self, self.arena, self.training, bns, n_layers, n_heads, We_all,
                                               *[p.detach() for p in t])
                    self.arena = None
                self.a_ext_t, self.values = out[0], list(out[1:])
                return self.values
            self.A, self.t = self._coefficients(self.projs, self.training)       # [L, d, Cm] (zero beyond C_l), [L, d]
            Wv = torch.cat([a.value_proj.weight for a in attns], 0).to(md)                          # [N0, d]
            bv = torch.cat([a.value_proj.bias for a in attns], 0)
            Aext = torch.cat([self.A, self.t.unsqueeze(-1)], -1)                                    # [L, d, Cm + 1]
            Fv = _fold_matmul(Wv, Aext)
            with torch.no_grad():
                Fe = _fold_matmul(We_all, Aext.detach())
            if grad and (Fv.requires_grad or bv.requires_grad or any(x.requires_grad for x in self.xs)):
                vals = _TokProjectFn.apply(self, self.arena, n_layers, n_heads, Fv, bv, Fe, *self.xs)
            else:
                vals = _TokProjectFn.forward(_NoCtx(), self, self.arena, n_layers, n_heads, Fv.detach(), bv.detach(), Fe,
                                             *self.xs)
                self.arena = None
        self.values = list(vals)
        return self.values

    def rank(self, valid_u8):
        """max over classes of enc_score_head(LayerNorm(enc_output.0(valid * feats))) for every token -> [B, Lv] fp32."""
        if self.scores is not None:
            if valid_u8 is not self.valid_u8 and valid_u8.data_ptr() != self.valid_u8.data_ptr():
                raise RuntimeError("tamtr_b200: FoldedTokens.rank() called with another validity mask than project()")
            return self.scores
        rk = self.rank_consts
        out = torch.empty(self.B, self.Lv, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().tamtr_rank_tokens(self.E.data_ptr(), self.raw.data_ptr(), rk["eb"].data_ptr(), valid_u8.data_ptr(),
                                              rk["bw"].data_ptr(), rk["sw"].data_ptr(), rk["ck"].data_ptr(), out.data_ptr(),
                                              _lib.dtype_code(self.E), self.B, self.Lv, self.d, rk["nc"], rk["npad"],
                                              float(rk["eps"]), _lib.stream_ptr(self.device))
        _lib.check(rc, "rank_tokens")
        return out

    # --- sparse side: the selected tokens, differentiably
    def rows(self, flat_idx):
        """feats.reshape(-1, d)[flat_idx] (flat_idx = image * Lv + token) recomputed from X through A and t."""
        md = MATH_DTYPE
        if self.a_ext_t is not None:        # fused glue: one gather kernel + one product
            L, K, R = len(self.xs), _kpad(self.Cm), flat_idx.numel()
            xcat = torch.empty(R, L * K, dtype=torch.float32, device=self.device)
            idx = flat_idx.contiguous().long()
            with torch.cuda.device(self.device):
                rc = _lib.lib().tamtr_fold_gather(L, self.Lv, _iarr(self.cs), _iarr(self.starts), _iarr(self.hw),
                                                  _parr(self.xs), idx.data_ptr(), xcat.data_ptr(), R,
                                                  _lib.stream_ptr(self.device))
            _lib.check(rc, "fold_gather")
            with torch.autocast(self.device.type, enabled=False):
                if torch.is_grad_enabled() and self.a_ext_t.requires_grad:
                    return _RowsFn.apply(self, idx, xcat, self.a_ext_t)
                return _mm_tf32(xcat, self.a_ext_t.reshape(L * K, -1))
        img = torch.div(flat_idx, self.Lv, rounding_mode="floor")
        tok = flat_idx - img * self.Lv
        out = None
        with torch.autocast(self.device.type, enabled=False):
            for l, x in enumerate(self.xs):
                C = self.cs[l]
                rel = tok - self.starts[l]
                inside = ((rel >= 0) & (rel < self.hw[l])).to(md).unsqueeze(-1)
                xr = x.flatten(2)[img, :, rel.clamp(0, self.hw[l] - 1)].to(md)                      # [R, C_l]
                y = (torch.addmm(self.t[l], xr, self.A[l, :, :C].t())) * inside
                out = y if out is None else out + y
        return out


class _NoCtx:
    """Stand-in for the autograd context when the projection runs without a graph (inference)."""
    needs_input_grad = ()

    def save_for_backward(self, *a):
        pass

    def set_materialize_grads(self, v):
        pass


def _rank_constants_fused(enc_linear, enc_norm, score_linear, fused):
    """_rank_constants + the ranking operand in one kernel (tamtr_fold_rank_consts); None when the parameters do not fit it."""
    We, eb, sw, sb = enc_linear.weight, enc_linear.bias, score_linear.weight, score_linear.bias
    lw, lb = enc_norm.weight, enc_norm.bias
    lin = [We, eb, sw, sb]
    if eb is None or sb is None or lw is None or lb is None or not all(t.is_cuda and t.is_contiguous() for t in lin + [lw, lb]):
        return None
    if any(t.dtype != We.dtype for t in lin) or We.dtype not in (torch.float32, torch.bfloat16) \
            or lw.dtype != torch.float32 or lb.dtype != torch.float32:
        return None
    nc, d = sw.shape
    fused = fused and nc + 1 <= 64
    npad = (nc + (1 if fused else 0) + 15) // 16 * 16
    if not fused:
        return None             # (tamtr_rank_tokens wants its constants as separate vectors: the torch path builds them)
    we_all = torch.empty(d + npad, d, dtype=torch.float32, device=We.device)
    consts = torch.empty(2 + 3 * npad, dtype=torch.float32, device=We.device)
    with torch.cuda.device(We.device):
        rc = _lib.lib().tamtr_fold_rank_consts(We.data_ptr(), eb.data_ptr(), sw.data_ptr(), sb.data_ptr(), lw.data_ptr(),
                                               lb.data_ptr(), we_all.data_ptr(), consts.data_ptr(), d, nc, npad, 1,
                                               _lib.dtype_code(We), _lib.stream_ptr(We.device))
    _lib.check(rc, "fold_rank_consts")
    we_all[d:] = _mm_tf32(we_all[d:], we_all[:d])          # tail rows: coefficients @ We
    return {"We_all": we_all, "nc": nc, "npad": npad, "eps": enc_norm.eps, "fused": True, "consts": consts}


def _rank_constants(enc_linear, enc_norm, score_linear, md, fused=False):
    """Constants of the ranking (include/tamtr_b200.h, tamtr_rank_tokens / tamtr_tok_project_rank): the score head folded
    with the LayerNorm weight.  Wr [npad, d]: rows of the projection's tail -- class k < nc: (score.weight * ln.weight)[k] @
    enc_output.0.weight; fused mode: the last row is enc_bias @ enc_output.0.weight (E . enc_bias, for the statistics)."""
    Wp = score_linear.weight.to(md) * enc_norm.weight.to(md)                                        # [nc, d]
    nc, d = Wp.shape
    fused = fused and nc + 1 <= 64
    npad = (nc + (1 if fused else 0) + 15) // 16 * 16
    We, eb = enc_linear.weight.to(md), enc_linear.bias.to(md)
    Wpad = torch.zeros(npad, d, dtype=md, device=Wp.device)
    Wpad[:nc] = Wp
    Wr = Wpad @ We
    bw, sw = Wp @ eb, Wp.sum(1)
    ck = score_linear.weight.to(md) @ enc_norm.bias.to(md) + score_linear.bias.to(md)
    out = {"Wr": Wr, "nc": nc, "npad": npad, "eps": enc_norm.eps, "fused": fused}
    if fused:
        Wr[npad - 1] = eb @ We
        consts = torch.zeros(2 + 3 * npad, dtype=md, device=Wp.device)
        consts[0], consts[1] = eb.sum(), (eb * eb).sum()
        consts[2:2 + nc], consts[2 + npad:2 + npad + nc], consts[2 + 2 * npad:2 + 2 * npad + nc] = bw, sw, ck
        out["consts"] = (consts if md != torch.float32 else consts.float()).contiguous()
    else:
        out.update({"eb": eb.float().contiguous(), "bw": bw.float().contiguous(), "sw": sw.float().contiguous(),
                    "ck": ck.float().contiguous()})
    return out
