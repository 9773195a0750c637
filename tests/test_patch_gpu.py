"""patch.enable() binds our forward() methods onto the REFERENCE's classes, so they may only use what the reference's
__init__ creates.  /root/reference does not travel to the GPU box; these stand-ins are built attribute-for-attribute like
ultralytics/nn/modules/transformer.py:501-525 and ultralytics/nn/extra_modules/block.py:197-206 (no helper methods of
ours), get our forward bound the way patch.py does it, and must reproduce our own modules on the GPU."""
import pytest
import torch
import torch.nn as nn

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu


class _RefConv(nn.Module):                      # ultralytics Conv(c1, c2, k, act=False): conv / bn / act
    def __init__(self, c1, c2, k):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, 1, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


def test_decoder_layer_forward_runs_on_a_reference_shaped_instance(cuda_lib):
    from tamtr_b200 import modules

    class RefLayer(nn.Module):
        def __init__(self, d_model=256, n_heads=8, d_ffn=512, dropout=0., act=nn.ReLU(), n_levels=3, n_points=4):
            super().__init__()
            self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
            self.dropout1 = nn.Dropout(dropout)
            self.norm1 = nn.LayerNorm(d_model)
            self.cross_attn = modules.MSDeformAttn(d_model, n_levels, n_heads, n_points)
            self.dropout2 = nn.Dropout(dropout)
            self.norm2 = nn.LayerNorm(d_model)
            self.linear1 = nn.Linear(d_model, d_ffn)
            self.act = act
            self.dropout3 = nn.Dropout(dropout)
            self.linear2 = nn.Linear(d_ffn, d_model)
            self.dropout4 = nn.Dropout(dropout)
            self.norm3 = nn.LayerNorm(d_model)

        @staticmethod
        def with_pos_embed(tensor, pos):
            return tensor if pos is None else tensor + pos

    RefLayer.forward = modules.DeformableTransformerDecoderLayer.forward
    ours = modules.DeformableTransformerDecoderLayer(256, 8, 512, 0.0, nn.ReLU(), 3, 4)
    seeding.seeded_fill(ours, 5)
    ref_like = RefLayer()
    ref_like.load_state_dict(ours.state_dict(), strict=True)
    ours.cuda()
    ref_like.cuda()
    shapes = [[20, 20], [10, 10], [5, 5]]
    B, Lq = 2, 40
    embed = seeding.seeded_tensor(6, "e", (B, Lq, 256)).cuda()
    feats = seeding.seeded_tensor(6, "f", (B, 525, 256)).cuda()
    refer = seeding.seeded_uniform(6, "r", (B, Lq, 4), 0.2, 0.8).cuda()
    pos = seeding.seeded_tensor(6, "p", (B, Lq, 256)).cuda()
    mask = (torch.rand(Lq, Lq) < 0.1).cuda()
    outs = []
    for m in (ours, ref_like):
        e = embed.clone().requires_grad_()
        o = m(e, refer, feats, shapes, None, mask, pos)
        o.square().sum().backward()
        outs.append((o.detach(), e.grad, m.norm3.weight.grad))
    for a, b in zip(*outs):
        assert rel_l2(a, b) < 1e-6


def test_max_sigmoid_block_forward_runs_on_a_reference_shaped_instance(cuda_lib):
    from tamtr_b200 import modules

    class RefBlock(nn.Module):
        def __init__(self, c1, c2, nh=1, ec=128, gc=512, scale=False):
            super().__init__()
            self.nh = nh
            self.hc = c2 // nh
            self.ec = _RefConv(c1, ec, 1) if c1 != ec else None
            self.gl = nn.Linear(gc, ec)
            self.bias = nn.Parameter(torch.zeros(nh))
            self.proj_conv = _RefConv(c1, c2, 3)
            self.scale = nn.Parameter(torch.ones(1, nh, 1, 1)) if scale else 1.0

    RefBlock.forward = modules.MaxSigmoidAttnBlock.forward
    ours = modules.MaxSigmoidAttnBlock(128, 128, nh=4, ec=128)
    seeding.seeded_fill(ours, 9)
    ref_like = RefBlock(128, 128, nh=4, ec=128)
    ref_like.load_state_dict(ours.state_dict(), strict=True)
    x = seeding.seeded_tensor(10, "x", (2, 128, 40, 40)).bfloat16().cuda()
    guide = seeding.seeded_tensor(10, "g", (2, 10, 512)).bfloat16().cuda()
    ours = ours.bfloat16().cuda().eval()
    ref_like = ref_like.bfloat16().cuda().eval()
    with torch.no_grad():                      # inference: the fused tcgen05 conv path with the folded BatchNorm
        a, b = ours(x, guide), ref_like(x, guide)
    assert torch.equal(a, b)
    ours.train()
    ref_like.train()
    a, b = ours(x, guide), ref_like(x, guide)  # training: tensor-core conv + BatchNorm statistics in torch
    assert torch.equal(a, b)


from refshape import RefMEH as _RefMEH, install_like_enable      # noqa: E402  (reference-shaped stand-ins)


def test_whole_head_through_the_enable_rebinding(cuda_lib):
    """What patch.enable() binds onto the reference's ManbaWorldDecoder / TextDeformableTransformerDecoder / VSSBlock
    (tested against the real classes on the CPU box: tests/test_patch.py) reproduces tamtr_b200.head.ManbaWorldDecoder
    bit for bit on the GPU -- train step with a denoising group and eval -- and launches the same kernels."""
    from tamtr_b200 import head
    install_like_enable()

    torch.manual_seed(0)
    ours = head.ManbaWorldDecoder(10, [64, 128, 256], 256, 30, 4, 8, 2, dims=[64, 128, 256], drop_path=(0.0, 0.0, 0.0))
    seeding.seeded_fill(ours, 21)
    with torch.no_grad():                       # keep the seeded VSS parameters in a sane range for the recurrence
        for blk in ours.VSSBlocks:
            blk.op.A_logs.copy_(torch.log(torch.arange(1, 17, dtype=torch.float32)).repeat(blk.op.A_logs.shape[0], 1))
            blk.op.dt_projs_bias.fill_(-3.0)
    ours.cuda().train()
    ref_like = _RefMEH(ours)
    assert not hasattr(ref_like, "_encode") or "_encode" in _RefMEH.__dict__      # nothing of ours beyond install()
    ref_like.cuda().train()
    with torch.no_grad():                       # the stand-in's own copies of the stacked SS2D parameters
        for rb, ob in zip(ref_like.VSSBlocks, ours.VSSBlocks):
            for name in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
                getattr(rb.op, name).copy_(getattr(ob.op, name))
    B = 2
    xs = [seeding.seeded_smooth_map(30 + i, "x", (B, c, s, s)).cuda() for i, (c, s) in enumerate(((64, 32), (128, 16), (256, 8)))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(31, "t", (B, 10, 512)), dim=-1).cuda()
    g = torch.Generator().manual_seed(4)
    groups = [5, 3]
    batch = {"cls": torch.randint(0, 10, (8,), generator=g), "bboxes": torch.cat(
        [0.2 + 0.6 * torch.rand(8, 2, generator=g), 0.05 + 0.2 * torch.rand(8, 2, generator=g)], -1),
        "batch_idx": torch.tensor([0] * 5 + [1] * 3), "gt_groups": groups}
    torch.manual_seed(11)
    plan = ours.plan_cdn(batch)
    outs, launches = [], []
    for m in (ours, ref_like):
        before = cuda_lib.launch_count()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m([x.clone() for x in xs], text, plan.to("cuda"))
        (out[0].float().square().mean() + out[1].float().sigmoid().mean()).backward()
        torch.cuda.synchronize()
        launches.append(cuda_lib.launch_count() - before)
        outs.append(out)
    assert launches[0] == launches[1] and launches[0] > 40, launches
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert torch.equal(a, b)
    ours.eval()
    ref_like.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ya, yb = ours(xs, text)[0], ref_like(xs, text)[0]
    assert torch.equal(ya, yb)


def test_scan_extension_shim_runs_the_reference_autograd_function(cuda_lib):
    """csms6s.py:252-270 (SelectiveScanCore) calls selective_scan_cuda_core.fwd / .bwd; enable() installs
    vss.ScanExtensionShim under that name.  The same call sequence here, against the recurrence oracle."""
    from oracle import vss_ref
    from tamtr_b200 import vss
    ext = vss.ScanExtensionShim
    u, dt = seeding.seeded_tensor(2, "u", (2, 128, 50)), seeding.seeded_tensor(2, "dt", (2, 128, 50)) - 2.0
    A = -(0.5 + 15.0 * seeding.seeded_uniform(2, "A", (128, 16)))
    Bm, Cm = seeding.seeded_tensor(2, "B", (2, 4, 16, 50)), seeding.seeded_tensor(2, "C", (2, 4, 16, 50))
    D, bias = seeding.seeded_tensor(2, "D", (128,)), seeding.seeded_tensor(2, "b", (128,)) * 0.1
    dout = seeding.seeded_tensor(2, "g", (2, 128, 50))
    cu = [t.cuda() for t in (u, dt, A, Bm, Cm, D, bias)]
    out, x, *rest = ext.fwd(*cu, True, 1)
    du, ddelta, dA, dB, dC, dD, dbias, *rest = ext.bwd(*cu, dout.cuda(), x, True, 1)
    leaves = [t.clone().double().requires_grad_() for t in (u, dt, A, Bm, Cm, D, bias)]
    ref = vss_ref.selective_scan(*leaves)
    ref.backward(dout.double())
    assert rel_l2(out, ref) < 1e-5
    for got, leaf in zip((du, ddelta, dA, dB, dC, dD, dbias), leaves):
        assert rel_l2(got, leaf.grad) < 1e-4


def test_fused_max_sigmoid_block_after_model_fuse(cuda_lib):
    """model.fuse() (nn/tasks.py:131-136) folds proj_conv's BatchNorm into its Conv2d (which gains a bias) and deletes
    the `bn` attribute: the patched forward must keep working and agree with the unfused block."""
    from tamtr_b200 import modules
    blk = modules.MaxSigmoidAttnBlock(128, 128, nh=4, ec=128)
    seeding.seeded_fill(blk, 12)
    blk = blk.cuda().eval()
    x = seeding.seeded_tensor(13, "x", (2, 128, 20, 20)).bfloat16().cuda()
    guide = seeding.seeded_tensor(13, "g", (2, 10, 512)).cuda()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        want = blk(x, guide).float()
        pc = blk.proj_conv                      # what ultralytics.utils.torch_utils.fuse_conv_and_bn produces
        s = pc.bn.weight / torch.sqrt(pc.bn.running_var + pc.bn.eps)
        fused = nn.Conv2d(128, 128, 3, 1, 1, bias=True).cuda()
        fused.weight.copy_(pc.conv.weight * s.view(-1, 1, 1, 1))
        fused.bias.copy_(pc.bn.bias - pc.bn.running_mean * s)
        pc.conv = fused
        delattr(pc, "bn")
        pc.forward = lambda t: pc.act(pc.conv(t))
        got = blk(x, guide).float()
    assert rel_l2(got, want) < 2e-2
