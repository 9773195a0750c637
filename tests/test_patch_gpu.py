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
