"""GPU: the fused decoder self-attention (csrc/selfattn.cu) against softmax(q k^T / sqrt(Dh) + mask) v evaluated in fp64 on the
same bf16-rounded projections (what nn.MultiheadAttention computes after its input projection, transformer.py:544-548), forward
and all gradients; bf16 bar of the north star: <= 2e-2 relative."""
import pytest
import torch

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu


def _reference(qk, v, blocked, H):
    B, L, d2 = qk.shape
    d = d2 // 2
    q = qk[..., :d].view(B, L, H, d // H).transpose(1, 2)
    k = qk[..., d:].view(B, L, H, d // H).transpose(1, 2)
    vv = v.view(B, L, H, d // H).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / (d // H) ** 0.5
    if blocked is not None:
        s = s.masked_fill(blocked.bool(), float("-inf"))
    return (torch.softmax(s, -1) @ vv).transpose(1, 2).reshape(B, L, d)


def _cdn_mask(L, n_dn, group):
    """models/utils/ops.py:273-284: matching queries do not see denoising queries, denoising groups do not see each other."""
    m = torch.zeros(L, L, dtype=torch.bool)
    m[n_dn:, :n_dn] = True
    for g0 in range(0, n_dn, group):
        m[g0:g0 + group, :g0] = True
        m[g0:g0 + group, g0 + group:n_dn] = True
    return m


@pytest.mark.parametrize("B,L,H,Dh,mask", [(2, 300, 8, 64, "cdn"), (1, 77, 4, 32, None), (2, 130, 2, 64, "random"),
                                          (1, 64, 8, 32, "cdn"), (1, 513, 1, 64, "random")])
def test_self_attention_matches_reference(cuda_lib, B, L, H, Dh, mask):
    from tamtr_b200 import ops
    d = H * Dh
    qk = (seeding.seeded_tensor(L, "qk", (B, L, 2 * d)) * 1.5).bfloat16()
    v = seeding.seeded_tensor(L, "v", (B, L, d)).bfloat16()
    go = seeding.seeded_tensor(L, "go", (B, L, d)).bfloat16()
    if mask == "cdn":
        n_dn = (2 * L // 3) // 8 * 8
        blocked = _cdn_mask(L, n_dn, n_dn // 4)
    elif mask == "random":
        blocked = seeding.seeded_uniform(L, "m", (L, L)) > 0.6
        blocked.fill_diagonal_(False)                      # every query sees at least itself
    else:
        blocked = None
    qr, vr = qk.double().requires_grad_(), v.double().requires_grad_()
    want = _reference(qr, vr, blocked, H)
    want.backward(go.double())
    qc, vc = qk.cuda().requires_grad_(), v.cuda().requires_grad_()
    before = cuda_lib.launch_count()
    bits = ops.attention_mask_bits(None if blocked is None else blocked.cuda())
    before = cuda_lib.launch_count()
    got = ops._SelfAttnFn.apply(qc, vc, bits, H)
    got.backward(go.cuda())
    torch.cuda.synchronize()
    assert cuda_lib.launch_count() - before == 3                     # one forward, two backward kernels
    assert got.dtype == torch.bfloat16 and rel_l2(got, want) < 1e-2, rel_l2(got, want)
    assert rel_l2(vc.grad, vr.grad) < 2e-2, rel_l2(vc.grad, vr.grad)
    assert rel_l2(qc.grad, qr.grad) < 2e-2, rel_l2(qc.grad, qr.grad)


def test_module_level_equals_library_attention(cuda_lib):
    """ops.self_attention (packed in-projection + fused attention + out-projection) against the same function on the
    library's scaled_dot_product_attention, under bf16 autocast, with the denoising mask: outputs and parameter gradients."""
    from tamtr_b200 import ops
    torch.manual_seed(5)
    mha = torch.nn.MultiheadAttention(512, 8, dropout=0.0).cuda()
    x = seeding.seeded_tensor(3, "x", (2, 300, 512)).cuda()
    pos = seeding.seeded_tensor(3, "pos", (2, 300, 512)).cuda()
    mask = _cdn_mask(300, 200, 40).cuda()
    probe = seeding.seeded_tensor(3, "p", (2, 300, 512)).cuda()

    def run(fused):
        ops.FUSED_SELF_ATTENTION = fused
        try:
            mha.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = ops.self_attention(mha, xi + pos, xi, mask)
            (y.float() * probe).sum().backward()
            return y.float().detach(), xi.grad, {k: p.grad.clone() for k, p in mha.named_parameters()}
        finally:
            ops.FUSED_SELF_ATTENTION = True
    y1, g1, p1 = run(True)
    y0, g0, p0 = run(False)
    assert rel_l2(y1, y0) < 2e-2 and rel_l2(g1, g0) < 2e-2, (rel_l2(y1, y0), rel_l2(g1, g0))
    for k in p0:
        assert rel_l2(p1[k], p0[k]) < 3e-2, (k, rel_l2(p1[k], p0[k]))
