"""GPU: the selective-scan kernels (csrc/sscan.cu) against the oracle restatement of the published recurrence
(oracle/vss_ref.selective_scan + autograd), and the product VSSBlock against the UNMODIFIED reference VSSBlock
(tests/golden/vss.pt).  Tolerances: fp32 <= 1e-4 relative on outputs (north_star); the gradients that are sums over
thousands of positions / channels (dA, dB, dC, d bias) <= 1e-3."""
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import seeding, vss_ref

pytestmark = pytest.mark.gpu


def _scan_inputs(seed, b, k, d, l, big_dt=False):
    n = 16
    u = seeding.seeded_tensor(seed, "u", (b, k * d, l))
    delta = seeding.seeded_tensor(seed, "dt", (b, k * d, l)) * (8.0 if big_dt else 1.0) - (0.0 if big_dt else 2.0)
    A = -(0.5 + 15.5 * seeding.seeded_uniform(seed, "A", (k * d, n)))
    B = seeding.seeded_tensor(seed, "B", (b, k, n, l))
    C = seeding.seeded_tensor(seed, "C", (b, k, n, l))
    D = 1.0 + 0.2 * seeding.seeded_tensor(seed, "D", (k * d,))
    bias = seeding.seeded_tensor(seed, "bias", (k * d,)) - 3.0
    return [u, delta, A, B, C, D, bias]


@pytest.mark.parametrize("b,k,d,l,big", [(2, 4, 128, 70, False), (1, 2, 256, 257, False), (1, 4, 128, 64, True),
                                         (3, 1, 128, 1, False), (1, 4, 128, 33, False)])
def test_selective_scan_matches_oracle(cuda_lib, b, k, d, l, big):
    from tamtr_b200.vss import selective_scan
    ins = _scan_inputs(l, b, k, d, l, big)
    ref_in = [t.clone().requires_grad_() for t in ins]
    y_ref = vss_ref.selective_scan(*ref_in, True)
    probe = seeding.seeded_tensor(l, "p", y_ref.shape)
    (y_ref * probe).sum().backward()
    cu = [t.cuda().requires_grad_() for t in ins]
    y = selective_scan(*cu, True)
    (y * probe.cuda()).sum().backward()
    assert rel_l2(y, y_ref) < 1e-4, rel_l2(y, y_ref)
    names = ["u", "delta", "A", "B", "C", "D", "bias"]
    for name, a, r in zip(names, cu, ref_in):
        tol = 1e-4 if name in ("u", "delta") else 1e-3
        assert rel_l2(a.grad, r.grad) < tol, (name, rel_l2(a.grad, r.grad))
    with torch.no_grad():                                            # inference: no checkpoints, same values
        assert torch.equal(selective_scan(*[t.detach() for t in cu], True), y.detach())


@pytest.mark.parametrize("name", ["c128_12x16", "c256_9x9", "c512_8x10"])
def test_vss_block_matches_reference(cuda_lib, name):
    from tamtr_b200.vss import VSSBlock
    gold = load_golden("vss")["cases"][name]
    c, b, h, w = gold["shape"]
    blk = VSSBlock(hidden_dim=c, drop_path=0.0)
    assert {k: tuple(v.shape) for k, v in blk.state_dict().items()} == gold["manifest"]
    vss_ref.seed_block(blk, gold["param_seed"])
    blk.cuda()
    x = seeding.seeded_tensor(600 + c, "x", (b, h, w, c)).cuda().requires_grad_()
    probe = seeding.seeded_tensor(600 + c, "probe", (b, h, w, c)).cuda()
    y = blk(x)
    (y * probe).sum().backward()
    assert rel_l2(y, gold["y"]) < 1e-4 and rel_l2(x.grad, gold["grad_x"]) < 1e-4
    for k, p in blk.named_parameters():
        n = gold["grad_param_norms"][k]
        assert abs(p.grad.double().norm().item() - n) < 1e-3 * max(n, 1e-6), k
    assert rel_l2(blk.op.A_logs.grad, gold["grad_A_logs"]) < 1e-3
    assert rel_l2(blk.op.x_proj_weight.grad, gold["grad_x_proj"]) < 1e-3
    assert rel_l2(blk.op.dt_projs_bias.grad, gold["grad_dt_bias"]) < 1e-3
    blk.train()
    blk.drop_path.drop_prob = 0.1                                    # head.py:1097: stochastic depth in training
    torch.manual_seed(0)
    out = blk(x.detach())
    assert out.shape == x.shape and torch.isfinite(out).all()


def test_bf16_inputs_are_converted_on_load(cuda_lib):
    """u / delta in bf16 (what SS2D hands over under autocast): the kernel converts on load, so the result is bit-identical
    to the reference's order of operations -- cast to fp32 first (vmamba.py:985-986), then scan -- and the input
    gradients are that path's gradients rounded once to bf16."""
    from tamtr_b200.vss import selective_scan
    ins = _scan_inputs(5, 2, 4, 128, 96)
    u16, d16 = ins[0].bfloat16().cuda(), ins[1].bfloat16().cuda()
    rest = [t.cuda() for t in ins[2:]]
    a = [u16.clone().requires_grad_(), d16.clone().requires_grad_()] + [t.clone().requires_grad_() for t in rest]
    b = [u16.float().requires_grad_(), d16.float().requires_grad_()] + [t.clone().requires_grad_() for t in rest]
    ya, yb = selective_scan(*a, True), selective_scan(*b, True)
    assert ya.dtype == torch.float32 and torch.equal(ya, yb)
    probe = seeding.seeded_tensor(5, "p", ya.shape).cuda()
    (ya * probe).sum().backward()
    (yb * probe).sum().backward()
    assert a[0].grad.dtype == torch.bfloat16
    assert torch.equal(a[0].grad, b[0].grad.bfloat16()) and torch.equal(a[1].grad, b[1].grad.bfloat16())
    for x, y in zip(a[2:], b[2:]):
        assert rel_l2(x.grad, y.grad) < 1e-5                         # fp32 reductions with atomics: order may differ
    odd = _scan_inputs(6, 1, 4, 128, 81)                             # odd L: rows are not 4-byte aligned -> fp32 path
    y_odd = selective_scan(odd[0].bfloat16().cuda(), odd[1].bfloat16().cuda(), *[t.cuda() for t in odd[2:]], True)
    ref = vss_ref.selective_scan(odd[0].bfloat16().float(), odd[1].bfloat16().float(), *odd[2:], True)
    assert rel_l2(y_odd, ref) < 1e-4


@pytest.mark.parametrize("b,d,h,w", [(2, 5, 12, 16), (1, 3, 33, 70), (2, 4, 64, 31), (1, 2, 1, 7)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cross_scan_and_merge_are_exact(cuda_lib, b, d, h, w, dtype):
    """csrc/crossscan.cu against the reference's flatten / transpose / flip / stack sequence (oracle/vss_ref.cross_scan,
    cross_merge = csms6s.py:6-13, 27-34): pure data movement and two-term sums in the reference's association, so
    EXACT in fp32 and in bf16 (sums rounded where the reference's tensor adds round), forward and backward."""
    from tamtr_b200.vss import cross_merge, cross_scan
    x = seeding.seeded_tensor(11, "x", (b, d, h, w)).to(dtype)
    xg = x.cuda().requires_grad_()
    xr = x.clone().requires_grad_()
    xs, xs_ref = cross_scan(xg), vss_ref.cross_scan(xr)
    assert xs.shape == (b, 4, d, h * w) and torch.equal(xs.cpu(), xs_ref)
    probe = seeding.seeded_tensor(12, "p", xs_ref.shape).to(dtype)
    xs.backward(probe.cuda())
    xs_ref.backward(probe)
    # CrossScan.backward is CrossMerge.forward (csms6s.py:17-24): same association, so exact as well
    assert torch.equal(xg.grad.cpu(), vss_ref.cross_merge(probe, h, w).view(b, d, h, w))
    if dtype == torch.float32:
        assert rel_l2(xg.grad, xr.grad) < 1e-6                        # autograd of the oracle sums in another order
    ys = seeding.seeded_tensor(13, "ys", (b, 4, d, h * w)).to(dtype)
    yg = ys.cuda().requires_grad_()
    y = cross_merge(yg, h, w)
    assert torch.equal(y.cpu(), vss_ref.cross_merge(ys, h, w))
    gp = seeding.seeded_tensor(14, "g", (b, d, h * w)).to(dtype)
    y.backward(gp.cuda())
    assert torch.equal(yg.grad.cpu(), vss_ref.cross_scan(gp.view(b, d, h, w)))


@pytest.mark.parametrize("b,d,h,w", [(2, 8, 12, 16), (1, 5, 33, 150), (2, 3, 40, 131), (1, 4, 1, 1)])
def test_dwconv3x3_silu_matches_library_conv(cuda_lib, b, d, h, w):
    """csrc/dwconv.cu against silu(conv2d(x)) as SS2D runs it (vmamba.py:1026-1027), evaluated by the library in fp64
    on the CPU: fp32 <= 1e-5 (forward, dx) / 1e-4 (weight / bias gradients: sums over b*h*w positions); bf16 activations
    <= 2e-2 against the same reference on bf16-rounded inputs."""
    import torch.nn as nn
    import torch.nn.functional as F
    from tamtr_b200.vss import dwconv3x3_silu
    conv = nn.Conv2d(d, d, 3, padding=1, groups=d)
    with torch.no_grad():
        conv.weight.copy_(seeding.seeded_tensor(21, "w", conv.weight.shape) * 0.4)
        conv.bias.copy_(seeding.seeded_tensor(21, "b", conv.bias.shape) * 0.3)
    x = seeding.seeded_tensor(22, "x", (b, d, h, w))
    g = seeding.seeded_tensor(23, "g", (b, d, h, w))
    for dtype, tol, tol_w in ((torch.float32, 1e-5, 1e-4), (torch.bfloat16, 2e-2, 2e-2)):
        xr = x.to(dtype).double().requires_grad_()
        gr = g.to(dtype).double()
        wr, br = conv.weight.detach().double().requires_grad_(), conv.bias.detach().double().requires_grad_()
        yr = F.silu(F.conv2d(xr, wr, br, padding=1, groups=d))
        yr.backward(gr)
        m = nn.Conv2d(d, d, 3, padding=1, groups=d).cuda()
        m.load_state_dict(conv.state_dict())
        xg = x.to(dtype).cuda().requires_grad_()
        y = dwconv3x3_silu(xg, m)
        y.backward(g.to(dtype).cuda())
        assert y.dtype == dtype and xg.grad.dtype == dtype and m.weight.grad.shape == conv.weight.shape
        assert rel_l2(y, yr) < tol and rel_l2(xg.grad, xr.grad) < tol
        assert rel_l2(m.weight.grad, wr.grad) < tol_w and rel_l2(m.bias.grad, br.grad) < tol_w


def test_config5_size_scan_and_block(cuda_lib):
    """BASELINE.json config 5 (1280x1280 input, batch 1 per GPU): the largest level is 320x320 = 102 400 positions.
    The scan against the oracle recurrence on one 32-channel group per direction, linearity in u at the full channel count
    (y(2u) - y(u) == y(u) - y(0): a size-independent property of the recurrence), and one VSSBlock forward + backward."""
    from tamtr_b200.vss import VSSBlock, selective_scan
    L = 320 * 320
    ins = _scan_inputs(31, 1, 4, 32, L)
    with torch.no_grad():
        y = selective_scan(*[t.cuda() for t in ins], True)
        ref = vss_ref.selective_scan(*ins, True)
    assert rel_l2(y, ref) < 1e-4
    u, rest = ins[0].cuda(), [t.cuda() for t in ins[1:]]
    with torch.no_grad():
        full = [seeding.seeded_tensor(32, "u", (1, 1024, L)).cuda(), (seeding.seeded_tensor(32, "dt", (1, 1024, L)) - 2.0).cuda(),
                -(0.5 + 15.5 * seeding.seeded_uniform(32, "A", (1024, 16))).cuda(),
                seeding.seeded_tensor(32, "B", (1, 4, 16, L)).cuda(), seeding.seeded_tensor(32, "C", (1, 4, 16, L)).cuda()]
        y1 = selective_scan(*full)
        y2 = selective_scan(2.0 * full[0], *full[1:])
        assert rel_l2(y2, 2.0 * y1) < 1e-5
    del y1, y2, full
    blk = VSSBlock(hidden_dim=128, drop_path=0.0).cuda()
    x = seeding.seeded_tensor(33, "x", (1, 320, 320, 128)).cuda().requires_grad_()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = blk(x)
    out.float().square().mean().backward()
    assert out.shape == x.shape and torch.isfinite(out).all() and torch.isfinite(x.grad).all()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in blk.parameters())


def test_scan_agrees_with_vllm_mamba_kernel(cuda_lib):
    """The extension the reference calls (selective_scan_cuda_core, VManba/csms6s.py:257) is not in its tree, so the scan has
    no reference output to pin against.  Closest independent implementation available in this image: vLLM's port of the
    mamba_ssm CUDA kernel (library code).  Forward only (vLLM ships no backward), in a subprocess (tests/scan_crosscheck_vllm.py);
    FAILS (does not skip) when vLLM cannot be imported or its kernel cannot run: the image ships it."""
    import json
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    try:
        proc = subprocess.run([sys.executable, os.path.join(here, "scan_crosscheck_vllm.py")], capture_output=True, text=True,
                              timeout=600)
    except subprocess.TimeoutExpired:
        pytest.fail("vLLM cross-check timed out")
    lines = [ln for ln in proc.stdout.splitlines() if ln.startswith("{")]
    if proc.returncode != 0 or not lines:
        pytest.fail("vLLM cross-check did not run: " + proc.stderr[-300:])
    res = json.loads(lines[-1])
    if "unavailable" in res:
        pytest.fail("vLLM selective_scan_fn unavailable (this image ships vLLM: the cross-check must run): "
                    + res["unavailable"])
    assert res["rel_l2"] and all(v < 1e-5 for v in res["rel_l2"].values()), res


@pytest.mark.parametrize("b,k,d,l", [(1, 4, 32, 4096), (1, 4, 64, 102400), (2, 4, 32, 9000), (1, 2, 32, 2050)])
def test_chunk_parallel_inference_scan(cuda_lib, b, k, d, l):
    """The chunk-parallel forward (state pass + output pass with folded carries) against the plain single-CTA scan of the
    same kernels -- mathematically the same recurrence, rounded differently where a piece's entry state is formed -- and
    against the oracle on the short case; bf16 inputs too."""
    from tamtr_b200 import vss
    lib = cuda_lib._lib.lib()
    pieces = lib.tamtr_selective_scan_chunks(b, k * d, l)
    assert pieces > 1 and lib.tamtr_selective_scan_chunks(16, 1024, 25600) == 1      # the training grid is left alone
    ins = [t.cuda() for t in _scan_inputs(51, b, k, d, l)]
    with torch.no_grad():
        before = cuda_lib.launch_count()
        y_chunked = vss.selective_scan(*ins, True)
        assert cuda_lib.launch_count() - before == 2
        vss.CHUNKED_INFERENCE = False
        try:
            y_plain = vss.selective_scan(*ins, True)
            y16_plain = vss.selective_scan(ins[0].bfloat16(), ins[1].bfloat16(), *ins[2:], True)
        finally:
            vss.CHUNKED_INFERENCE = True
        y16 = vss.selective_scan(ins[0].bfloat16(), ins[1].bfloat16(), *ins[2:], True)
    assert rel_l2(y_chunked, y_plain) < 2e-6 and rel_l2(y16, y16_plain) < 2e-6
    if l <= 4096:
        assert rel_l2(y_chunked, vss_ref.selective_scan(*[t.cpu() for t in ins], True)) < 1e-4
    # with gradients required the checkpointing forward is used (one launch), whatever the grid
    leaf = [t.clone().requires_grad_() for t in ins]
    before = cuda_lib.launch_count()
    vss.selective_scan(*leaf, True)
    assert cuda_lib.launch_count() - before == 1


@pytest.mark.parametrize("b,k,d,l", [(2, 4, 32, 48), (1, 4, 32, 95), (1, 2, 64, 130)])
def test_scan_backward_against_the_closed_form(cuda_lib, b, k, d, l):
    """The scan kernels' backward (all seven gradients) against an INDEPENDENT formulation: oracle/vss_ref's closed form
    (cumulative sums + masked contraction in fp64, autograd) -- not the position loop the recurrence oracle and the kernel
    share.  Lengths that are not multiples of the kernels' 16-position tiles included."""
    from oracle import vss_ref
    from tamtr_b200.vss import selective_scan
    n = 16
    mk = lambda name, shape: seeding.seeded_tensor(23, name, shape)
    base = [mk("u", (b, k * d, l)), mk("dt", (b, k * d, l)) - 1.5, -(0.5 + 12.0 * seeding.seeded_uniform(23, "A", (k * d, n))),
            mk("B", (b, k, n, l)), mk("C", (b, k, n, l)), 1.0 + 0.2 * mk("D", (k * d,)), 0.3 * mk("bias", (k * d,))]
    gout = mk("g", (b, k * d, l))
    ref = [t.clone().double().requires_grad_() for t in base]
    y_ref = vss_ref.selective_scan_closed_form(*ref)
    y_ref.backward(gout.double())
    ours = [t.clone().cuda().requires_grad_() for t in base]
    y = selective_scan(*ours)
    y.backward(gout.cuda())
    torch.cuda.synchronize()
    assert rel_l2(y, y_ref) < 1e-5
    for name, got, want in zip(("u", "delta", "A", "B", "C", "D", "delta_bias"), ours, ref):
        assert rel_l2(got.grad, want.grad) < 1e-4, (name, rel_l2(got.grad, want.grad))


@pytest.mark.parametrize("c,h,w,autocast", [(64, 12, 20, False), (128, 16, 16, True), (64, 7, 9, False)])
def test_fused_ss2d_equals_the_composition(cuda_lib, c, h, w, autocast):
    """The single-node SS2D (explicit backward, position-major GEMMs, fused out_norm + gate: vss._SS2DFn) against the
    op-by-op composition of the same kernels / library calls, outputs and every gradient; fp32 <= 1e-4, bf16 autocast <= 2e-2.
    (The composition itself is pinned to the unmodified reference VSSBlock by test_vss_block_matches_reference.)"""
    from tamtr_b200 import vss
    torch.manual_seed(3)
    blk = vss.VSSBlock(hidden_dim=c, drop_path=0.0).cuda()
    with torch.no_grad():
        for prm in blk.parameters():
            if prm.dim() == 1 and prm is not blk.op.Ds:
                prm.add_(0.1 * torch.randn_like(prm))
    x = seeding.seeded_tensor(9, "x", (2, h, w, c)).cuda()
    probe = seeding.seeded_tensor(9, "p", (2, h, w, c)).cuda()

    def run(fused):
        vss.FUSED_SS2D = fused
        try:
            blk.zero_grad(set_to_none=True)
            xi = x.clone().requires_grad_()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                y = blk(xi)
            (y.float() * probe).sum().backward()
            return y.detach().float(), xi.grad.clone(), {k: v.grad.clone() for k, v in blk.named_parameters()}
        finally:
            vss.FUSED_SS2D = True
    before = cuda_lib.launch_count()
    y1, gx1, gp1 = run(True)
    fused_launches = cuda_lib.launch_count() - before
    y0, gx0, gp0 = run(False)
    tol = 2e-2 if autocast else 1e-4
    assert fused_launches > 0 and rel_l2(y1, y0) < tol and rel_l2(gx1, gx0) < tol, (rel_l2(y1, y0), rel_l2(gx1, gx0))
    for name in gp0:
        assert gp1[name].shape == gp0[name].shape and gp1[name].dtype == gp0[name].dtype
        # sums over thousands of positions / channels in different orders (and, under autocast, bf16 partials)
        assert rel_l2(gp1[name], gp0[name]) < (5e-2 if autocast else 1e-3), (name, rel_l2(gp1[name], gp0[name]))


def test_colnorm_gate_matches_torch(cuda_lib):
    """LayerNorm over the channel dimension of a position-major tensor times silu(z) (csrc/vssfuse.cu) against
    F.layer_norm on the transposed tensor, forward values (fp32 <= 1e-5; bf16 gate and output <= 1e-2)."""
    import torch.nn.functional as F
    from tamtr_b200 import vss
    norm = torch.nn.LayerNorm(96).cuda()
    with torch.no_grad():
        norm.weight.add_(0.2 * torch.randn_like(norm.weight))
        norm.bias.add_(0.2 * torch.randn_like(norm.bias))
    y = (3.0 + 2.0 * seeding.seeded_tensor(4, "y", (2, 96, 301))).cuda()
    z = seeding.seeded_tensor(4, "z", (2, 96, 301)).cuda()
    want = (F.layer_norm(y.transpose(1, 2), (96,), norm.weight, norm.bias, norm.eps) * F.silu(z.transpose(1, 2))).transpose(1, 2)
    out, mean, rstd = vss.colnorm_gate(y, z, norm)
    assert rel_l2(out, want) < 1e-5 and rel_l2(mean, y.mean(1)) < 1e-6
    out16, _, _ = vss.colnorm_gate(y, z.bfloat16(), norm)
    assert out16.dtype == torch.bfloat16 and rel_l2(out16.float(), want) < 1e-2


def test_pyramid_levels_on_parallel_streams(cuda_lib):
    """head._apply_vss_blocks: the three levels' VSSBlocks issued on forked streams (parallel branches of a captured step)
    give what the sequential loop gives -- outputs bit-identical (no cross-level data), input gradients to rounding (the
    scan's dB / dC are fp32 atomics, whose order varies run to run anyway)."""
    from tamtr_b200 import head
    from tamtr_b200.vss import VSSBlock
    torch.manual_seed(11)
    blocks = torch.nn.ModuleList(VSSBlock(hidden_dim=c, drop_path=0.0) for c in (64, 128, 256)).cuda()
    xs = [seeding.seeded_tensor(31, f"x{i}", (2, c, s, s)).cuda() for i, (c, s) in enumerate(((64, 24), (128, 12), (256, 8)))]

    def run(parallel):
        head.VSS_PARALLEL_LEVELS = parallel
        try:
            leaves = [x.clone().requires_grad_() for x in xs]
            with torch.autocast("cuda", dtype=torch.bfloat16):
                outs = head._apply_vss_blocks(blocks, leaves)
            sum((o.float() ** 2).mean() for o in outs).backward()
            torch.cuda.synchronize()
            return [o.detach() for o in outs], [x.grad for x in leaves]
        finally:
            head.VSS_PARALLEL_LEVELS = True
    o1, g1 = run(True)
    o0, g0 = run(False)
    for a, b in zip(o1, o0):
        assert a.shape == b.shape and torch.equal(a, b)
    for a, b in zip(g1, g0):
        assert rel_l2(a, b) < 1e-3
