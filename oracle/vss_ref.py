"""TEST INFRASTRUCTURE ONLY -- CPU restatement of VMamba's VSSBlock / SS2D as TAM-TR's head uses it
(ultralytics/nn/modules/head.py:1092-1098,1134: VSSBlock(hidden_dim=dims[i], drop_path=0.1) on each pyramid level).

Follows (file:line under /root/reference/ultralytics/nn/extra_modules/VManba):
  VSSBlock._forward          vmamba.py:1236-1250   x + drop_path(op(norm(x))); x + drop_path(mlp(norm2(x)))
  SS2D.forwardv2             vmamba.py:1019-1038   in_proj, chunk, SiLU(z), depth-wise conv 3x3, SiLU, core, *z, out_proj
  SS2D.forward_corev2        vmamba.py:898-1017    cross scan, x_proj / dt_proj einsums, selective scan, cross merge, LayerNorm
  CrossScan / CrossMerge     csms6s.py:4-47        the four scan orders (row-major, column-major, and both reversed)
  Mlp                        vmamba.py:107-125

**parity unpinned at one boundary**: the selective scan itself lives in a third-party CUDA extension
(`selective_scan_cuda_core`, VMamba kernels/selective_scan; un-vendored and unpinned -- README.md:39-41 only links the
repository, it is not in requirements.txt) that is absent from /root/reference and from this image, and no reference
test pins it.  `selective_scan` below restates the published recurrence (Gu & Dao, "Mamba", 2023, eq. 2 with the
zero-order-hold discretisation of Alg. 2, as implemented by that extension's `selective_scan_ref`):
    delta_t = softplus(dt_t + bias)            (softplus(x) = x for x > 20)
    h_t     = exp(delta_t * A) * h_{t-1} + delta_t * B_t * u_t          (h_{-1} = 0, per channel d and state n)
    y_t     = <C_t, h_t> + D * u_t
(A forward-only cross-check of the CUDA scan against vLLM's port of the mamba_ssm kernel -- library code in this image, the
same kernel family -- lives in tests/test_vss_gpu.py; it narrows this gap, it does not close it.)
Everything AROUND the scan is pinned to the unmodified reference: `install_scan_extension()` plugs this function into the
reference's namespace under the missing extension's name so that the reference's own VSSBlock runs end to end
(oracle/make_goldens_vss.py -> tests/golden/vss.pt).
"""
import types

import torch
import torch.nn.functional as F


def selective_scan(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True):
    """u, delta [b, kd, l]; A [kd, n]; B, C [b, k, n, l]; D, delta_bias [kd] -> y [b, kd, l] (computed in fp32, or in
    fp64 when u is fp64; differentiable)."""
    b, kd, l = u.shape
    k, n = B.shape[1], A.shape[1]
    d = kd // k
    ft = torch.float64 if u.dtype == torch.float64 else torch.float32
    delta = delta.to(ft)
    if delta_bias is not None:
        delta = delta + delta_bias.to(ft).view(1, kd, 1)
    if delta_softplus:
        delta = torch.where(delta > 20.0, delta, F.softplus(delta, threshold=1e9))
    u32 = u.to(ft)
    Bx = B.to(ft).view(b, k, 1, n, l).expand(b, k, d, n, l).reshape(b, kd, n, l)
    Cx = C.to(ft).view(b, k, 1, n, l).expand(b, k, d, n, l).reshape(b, kd, n, l)
    h = u32.new_zeros(b, kd, n)
    ys = []
    for t in range(l):
        dt = delta[:, :, t].unsqueeze(-1)                                   # [b, kd, 1]
        h = torch.exp(dt * A.to(ft)) * h + dt * Bx[..., t] * u32[:, :, t].unsqueeze(-1)
        ys.append((h * Cx[..., t]).sum(-1))
    y = torch.stack(ys, -1)
    if D is not None:
        y = y + D.to(ft).view(1, kd, 1) * u32
    return y


def selective_scan_closed_form(u, delta, A, B, C, D=None, delta_bias=None):
    """The same map WITHOUT the recurrence: with S_t = sum_{r <= t} delta_r the state is
        h_t = sum_{s <= t} exp(A * (S_t - S_s)) * delta_s * B_s * u_s
    so  y_t = sum_n C_t[n] * sum_{s <= t} exp(A_n * (S_t - S_s)) * delta_s * B_s[n] * u_s + D * u_t,
    evaluated as one masked [l, l] kernel per (b, channel, state) in fp64 (every exponent is <= 0: no overflow).  An
    independent formulation -- a cumulative sum, an outer difference and a masked contraction instead of a loop over t --
    whose autograd graph shares nothing with `selective_scan` above: the second opinion for the CUDA kernels' backward
    (tests/test_vss_gpu.py) and for the recurrence restatement itself (tests/test_oracle_vss.py).  Small l only."""
    b, kd, l = u.shape
    k, n = B.shape[1], A.shape[1]
    d = kd // k
    f = torch.float64
    dl = delta.to(f)
    if delta_bias is not None:
        dl = dl + delta_bias.to(f).view(1, kd, 1)
    dl = torch.where(dl > 20.0, dl, torch.log1p(torch.exp(dl.clamp(max=20.0))))
    S = torch.cumsum(dl, -1)                                                     # [b, kd, l]
    mask = torch.ones(l, l, dtype=torch.bool, device=u.device).tril()
    diff = (S.unsqueeze(-1) - S.unsqueeze(-2)).masked_fill(~mask, 0.0)           # [b, kd, t, s] = S_t - S_s for s <= t
    decay = torch.exp(A.to(f).view(1, kd, n, 1, 1) * diff.unsqueeze(2)) * mask   # [b, kd, n, t, s]
    Bx = B.to(f).view(b, k, 1, n, l).expand(b, k, d, n, l).reshape(b, kd, n, l)
    Cx = C.to(f).view(b, k, 1, n, l).expand(b, k, d, n, l).reshape(b, kd, n, l)
    src = (dl * u.to(f)).unsqueeze(2) * Bx                                       # [b, kd, n, s]
    h = torch.einsum("bcnts,bcns->bcnt", decay, src)
    y = (h * Cx).sum(2)
    if D is not None:
        y = y + D.to(f).view(1, kd, 1) * u.to(f)
    return y


def install_scan_extension(csms6s_module):
    """Give the reference's csms6s.py the extension object it failed to import (csms6s.py:121-126): fwd/bwd with the
    extension's call signature (csms6s.py:257,266), implemented by the restatement above + autograd."""
    def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows):
        with torch.no_grad():
            out = selective_scan(u, delta, A, B, C, D, delta_bias, delta_softplus)
        return out, u.new_zeros(1)

    def bwd(u, delta, A, B, C, D, delta_bias, dout, x, delta_softplus, nrows):
        ins = [t.detach().clone().requires_grad_() for t in (u, delta, A, B, C, D, delta_bias)]
        with torch.enable_grad():
            out = selective_scan(*ins, delta_softplus)
        return torch.autograd.grad(out, ins, dout)
    csms6s_module.selective_scan_cuda_core = types.SimpleNamespace(fwd=fwd, bwd=bwd)


def cross_scan(x):
    """[b, d, h, w] -> [b, 4, d, h*w] (csms6s.py:6-13)."""
    a = x.flatten(2)
    c = x.transpose(2, 3).flatten(2)
    return torch.stack([a, c, a.flip(-1), c.flip(-1)], 1)


def cross_merge(ys, h, w):
    """[b, 4, d, h*w] -> [b, d, h*w] (csms6s.py:27-34)."""
    b, _, d, l = ys.shape
    s = ys[:, 0:2] + ys[:, 2:4].flip(-1)
    return s[:, 0] + s[:, 1].reshape(b, d, w, h).transpose(2, 3).reshape(b, d, l)


def ss2d(sd, p, x, scan=selective_scan):
    """SS2D.forwardv2 (forward_type "v2", channel-last input [b, h, w, c])."""
    b, h, w, _ = x.shape
    xz = F.linear(x, sd[p + ".in_proj.weight"])
    xi, z = xz.chunk(2, dim=-1)
    z = F.silu(z)
    xi = xi.permute(0, 3, 1, 2).contiguous()
    d = xi.shape[1]
    xi = F.silu(F.conv2d(xi, sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"], padding=1, groups=d))
    xw, dtw, dtb = sd[p + ".x_proj_weight"], sd[p + ".dt_projs_weight"], sd[p + ".dt_projs_bias"]
    k, r, n = xw.shape[0], dtw.shape[2], sd[p + ".A_logs"].shape[1]
    xs = cross_scan(xi)
    x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs, xw)
    dts, Bs, Cs = torch.split(x_dbl, [r, n, n], dim=2)
    dts = torch.einsum("b k r l, k d r -> b k d l", dts, dtw)
    ft = torch.float64 if x.dtype == torch.float64 else torch.float32        # vmamba.py:985-986 forces fp32 inputs
    ys = scan(xs.reshape(b, -1, h * w).to(ft), dts.reshape(b, -1, h * w).to(ft), -torch.exp(sd[p + ".A_logs"].to(ft)),
              Bs.contiguous().to(ft), Cs.contiguous().to(ft), sd[p + ".Ds"].to(ft), dtb.reshape(-1).to(ft), True)
    y = cross_merge(ys.view(b, k, -1, h * w), h, w)
    y = F.layer_norm(y.transpose(1, 2).contiguous(), (d,), sd[p + ".out_norm.weight"], sd[p + ".out_norm.bias"], 1e-5)
    y = y.view(b, h, w, d).to(x.dtype) * z
    return F.linear(y, sd[p + ".out_proj.weight"])


def vss_block(sd, p, x, scan=selective_scan):
    """VSSBlock._forward with drop_path inactive (eval mode or rate 0)."""
    c = x.shape[-1]
    x = x + ss2d(sd, p + ".op", F.layer_norm(x, (c,), sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-5), scan)
    y = F.layer_norm(x, (c,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])), sd[p + ".mlp.fc2.weight"],
                 sd[p + ".mlp.fc2.bias"])
    return x + y


def seed_block(block, seed):
    """Deterministic, key-named fill (independent of parameter registration order) that keeps the SSM parameters in
    their working range (A_logs = log of positive numbers, dt bias = inverse softplus of a small step)."""
    from oracle import seeding
    with torch.no_grad():
        for name, t in block.state_dict().items():
            if not t.is_floating_point():
                continue
            g = seeding._gen(seed, name)
            leaf = name.rsplit(".", 1)[-1]
            if leaf == "A_logs":
                t.copy_(torch.log(0.5 + 15.5 * torch.rand(t.shape, generator=g)))
            elif leaf == "dt_projs_bias":
                dt = torch.exp(torch.rand(t.shape, generator=g) * 4.6 - 6.9)
                t.copy_(dt + torch.log(-torch.expm1(-dt)))
            elif leaf == "Ds":
                t.copy_(1.0 + 0.2 * torch.randn(t.shape, generator=g))
            elif t.dim() >= 2:
                fan_in = t[0].numel()
                t.copy_(torch.randn(t.shape, generator=g) / fan_in ** 0.5)
            elif leaf == "weight":
                t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
            else:
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
