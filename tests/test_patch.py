"""enable()/disable() on the real reference package (only where /root/reference exists: the build container)."""
import copy
import io

import pytest
import torch

from oracle import reference_loader

pytestmark = pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present on this box")


def test_enable_rebinds_methods_in_place_and_disable_restores():
    import tamtr_b200
    ns = reference_loader.hot_path()
    T, U = ns.transformer, ns.utils
    orig_fn, orig_fwd = U.multi_scale_deformable_attn_pytorch, T.MSDeformAttn.forward
    cls_before = T.MSDeformAttn
    m = T.MSDeformAttn(64, 3, 4, 4)                     # built BEFORE enable()
    tamtr_b200.enable()
    try:
        assert tamtr_b200.enabled()
        assert T.MSDeformAttn is cls_before and isinstance(m, T.MSDeformAttn)
        assert T.multi_scale_deformable_attn_pytorch is tamtr_b200.ops.ms_deform_attn
        assert U.multi_scale_deformable_attn_pytorch is tamtr_b200.ops.ms_deform_attn
        assert T.MSDeformAttn.forward is not orig_fwd
        # deep copies, pickles and state_dicts are untouched by the patch
        m2 = copy.deepcopy(m)
        buf = io.BytesIO()
        torch.save(m, buf)
        buf.seek(0)
        m3 = torch.load(buf, weights_only=False)
        assert type(m2) is T.MSDeformAttn and type(m3) is T.MSDeformAttn
        assert list(m3.state_dict()) == list(m.state_dict())
        # no CPU fallback: the error carries the strings nn/tasks.py:256-264 looks for
        q = torch.zeros(1, 5, 64)
        with pytest.raises(RuntimeError) as e:
            m(q, torch.rand(1, 5, 1, 4), torch.zeros(1, 21, 64), [[4, 4], [2, 2], [1, 1]])
        assert "Not implemented on the CPU" in str(e.value) or "is_cuda" in str(e.value)
        head = ns.ContrastiveHeadMLP()
        with pytest.raises(RuntimeError):
            head(torch.zeros(1, 4, 128), torch.zeros(1, 3, 128))
    finally:
        tamtr_b200.disable()
    assert not tamtr_b200.enabled()
    assert T.MSDeformAttn.forward is orig_fwd and U.multi_scale_deformable_attn_pytorch is orig_fn
    out = m(torch.zeros(1, 5, 64), torch.rand(1, 5, 1, 4), torch.zeros(1, 21, 64), [[4, 4], [2, 2], [1, 1]])
    assert out.shape == (1, 5, 64)


def test_mirror_state_dicts_load_into_each_other():
    """Checkpoint compatibility both ways: same keys and shapes as the reference's modules."""
    from tamtr_b200.head import ManbaWorldDecoder, RTDETRDecoder
    from tamtr_b200.modules import MaxSigmoidAttnBlock
    ns = reference_loader.hot_path()
    ref = ns.RTDETRDecoder(nc=10, ch=(32, 64, 128), hd=64, nq=20, ndl=2)
    ours = RTDETRDecoder(nc=10, ch=(32, 64, 128), hd=64, nq=20, ndl=2)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(ours.state_dict(), strict=True)
    ref_b = ns.MaxSigmoidAttnBlock(64, 64, nh=2, ec=64)
    MaxSigmoidAttnBlock(64, 64, nh=2, ec=64).load_state_dict(ref_b.state_dict(), strict=True)
    ref_e = ns.MaxSigmoidAttnBlock(64, 64, nh=2, ec=32)      # with the optional embedding conv
    MaxSigmoidAttnBlock(64, 64, nh=2, ec=32).load_state_dict(ref_e.state_dict(), strict=True)
    # the whole head, VSSBlocks included: same keys and shapes both ways
    mine = ManbaWorldDecoder(10, [32, 64, 128], 64, 20, 4, 8, 2, dims=[32, 64, 128])
    ref_m = ns.ManbaWorldDecoder(10, [32, 64, 128], 64, 20, 4, 8, 2, dims=[32, 64, 128])
    assert {k: tuple(v.shape) for k, v in mine.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ref_m.state_dict().items()}
    mine.load_state_dict(ref_m.state_dict(), strict=True)
    ref_m.load_state_dict(mine.state_dict(), strict=True)
    # vss=False (the configuration of the head-level fixtures): VSSBlocks.* entries of a checkpoint are skipped
    ManbaWorldDecoder(10, [32, 64, 128], 64, 20, 4, 8, 2, vss=False).load_state_dict(ref_m.state_dict(), strict=True)


def test_enable_covers_the_decoupled_branch():
    """The cls / box samplers, their attention modules and the decoupled decoder layer (utils.py:92-191,
    transformer.py:300-495, 561-658) are rebound too, restored by disable(), and share state_dict keys with our mirrors."""
    import tamtr_b200
    from tamtr_b200 import modules
    ns = reference_loader.hot_path()
    T, U = ns.transformer, ns.utils
    before = (U.multi_scale_deformable_attn_pytorch_cls, T.multi_scale_deformable_attn_pytorch_box,
              T.MSDeformAttncls.forward, T.DecouplingDeformableTransformerDecoderLayer.forward)
    layer = T.DecouplingDeformableTransformerDecoderLayer(64, 4, 128, 0.0, torch.nn.ReLU(), 3, 4)
    ours = modules.DecouplingDeformableTransformerDecoderLayer(64, 4, 128, 0.0, torch.nn.ReLU(), 3, 4)
    ours.load_state_dict(layer.state_dict(), strict=True)
    layer.load_state_dict(ours.state_dict(), strict=True)
    tamtr_b200.enable()
    try:
        assert U.multi_scale_deformable_attn_pytorch_cls is tamtr_b200.ops.ms_deform_attn_cls
        assert T.multi_scale_deformable_attn_pytorch_cls is tamtr_b200.ops.ms_deform_attn_cls
        assert T.multi_scale_deformable_attn_pytorch_box is tamtr_b200.ops.ms_deform_attn_box
        assert T.MSDeformAttncls.forward is modules.MSDeformAttncls.forward
        assert T.MSDeformAttnbox.forward is modules.MSDeformAttnbox.forward
        x = torch.zeros(1, 5, 64)
        with pytest.raises(RuntimeError, match="is_cuda|Not implemented on the CPU"):
            layer(x, x, torch.rand(1, 5, 4), torch.zeros(1, 21, 64), [[4, 4], [2, 2], [1, 1]])
    finally:
        tamtr_b200.disable()
    assert (U.multi_scale_deformable_attn_pytorch_cls, T.multi_scale_deformable_attn_pytorch_box,
            T.MSDeformAttncls.forward, T.DecouplingDeformableTransformerDecoderLayer.forward) == before
    o_cls, o_box = layer(x, x, torch.rand(1, 5, 4), torch.zeros(1, 21, 64), [[4, 4], [2, 2], [1, 1]])
    assert o_cls.shape == o_box.shape == (1, 5, 64)
