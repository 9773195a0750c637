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


def test_enable_installs_the_whole_fast_path_on_the_real_model():
    """After enable() the model TAMTR.yaml builds runs the benchmarked path, not only its leaves: head glue, decoder,
    VSSBlocks, matcher and loss are rebound IN PLACE on the reference's own classes (nn/modules/head.py:1130-1264,
    transformer.py:850-891, VManba/vmamba.py:898-1038, models/utils/ops.py:48-121), and disable() restores everything."""
    import importlib
    import os
    import tamtr_b200
    from tamtr_b200 import head, loss, modules, patch, vss
    ns = reference_loader.hot_path()
    tasks = importlib.import_module("ultralytics.nn.tasks")
    vm = importlib.import_module("ultralytics.nn.extra_modules.VManba.vmamba")
    cs = importlib.import_module("ultralytics.nn.extra_modules.VManba.csms6s")
    lo = importlib.import_module("ultralytics.models.utils.loss")
    yaml = os.path.join(reference_loader.REFERENCE_ROOT, "ultralytics/cfg/models/TAMTR/TAMTR.yaml")
    model = tasks.RTDETRDetectionWorldModel(yaml, nc=10, verbose=False)          # built BEFORE enable()
    meh = model.model[-1]
    assert type(meh) is ns.ManbaWorldDecoder
    before = {c: dict(c.__dict__) for c in (ns.ManbaWorldDecoder, ns.RTDETRDecoder, ns.TextDeformableTransformerDecoder,
                                            vm.VSSBlock, vm.SS2D, ns.ops.HungarianMatcher, lo.DETRLoss,
                                            lo.RTDETRDetectionLoss, ns.TIAGELAN)}
    assert not hasattr(cs, "selective_scan_cuda_core")
    tamtr_b200.enable()
    try:
        for ref_cls, ours in ((ns.ManbaWorldDecoder, head.ManbaWorldDecoder), (ns.RTDETRDecoder, head.RTDETRDecoder)):
            for name in patch.HEAD_ATTRS:
                assert ref_cls.__dict__[name] is ours.__dict__.get(name, head._HeadBase.__dict__.get(name)), name
        for name in patch.DECODER_ATTRS:
            assert ns.TextDeformableTransformerDecoder.__dict__[name] is (
                modules.TextDeformableTransformerDecoder.__dict__.get(name, modules._DecoderBase.__dict__.get(name)))
        assert ns.TIAGELAN.forward is modules.TIAGELAN.forward and ns.TIAGELAN.forward_split is modules.TIAGELAN.forward
        assert vm.SS2D.forwardv2 is not before[vm.SS2D]["forwardv2"]
        assert vm.VSSBlock.forward is not before[vm.VSSBlock]["forward"]
        assert cs.selective_scan_cuda_core is vss.ScanExtensionShim
        assert ns.ops.HungarianMatcher.forward is loss.HungarianMatcher.forward
        assert lo.RTDETRDetectionLoss.forward is loss.RTDETRDetectionLoss.forward
        # the instance built before enable() is still the reference's class and picks the methods up
        assert type(meh) is ns.ManbaWorldDecoder and meh.forward.__func__ is head.ManbaWorldDecoder.forward
        assert type(meh.decoder).forward is modules.TextDeformableTransformerDecoder.forward
        assert all(vss._ss2d_supported(b.op) for b in meh.VSSBlocks)              # our SS2D path applies to what it builds
        # no CPU fallback anywhere on the path: the head refuses CPU maps with the message nn/tasks.py:256-264 expects
        xs = [torch.zeros(1, c, s, s) for c, s in zip((128, 256, 512), (16, 8, 4))]
        with pytest.raises(RuntimeError, match="is_cuda|Not implemented on the CPU"):
            meh.eval()(xs, torch.zeros(1, 10, 512))
        crit = lo.RTDETRDetectionLoss(nc=10, use_vfl=True)
        with pytest.raises(RuntimeError, match="is_cuda|Not implemented on the CPU"):
            crit((torch.zeros(2, 1, 5, 4), torch.zeros(2, 1, 5, 10)),
                 {"cls": torch.zeros(1, dtype=torch.long), "bboxes": torch.rand(1, 4), "gt_groups": [1]})
        # pickling / deep copies of the patched head keep working (the anchor cache stays out)
        copy.deepcopy(meh)
    finally:
        tamtr_b200.disable()
    for c, d in before.items():
        assert dict(c.__dict__) == d, c
    assert not hasattr(cs, "selective_scan_cuda_core")


def test_tiagelan_mirror_equals_reference_and_ignores_the_guide():
    """TIAGELAN (extra_modules/block.py:171-192) calls its MaxSigmoidAttnBlock and discards the result (:185): the output
    does not depend on the guide (SURVEY.md KAT #4) and the only trace of the call is the train-mode update of
    attn.proj_conv.bn's running statistics.  Our mirror / rebound forward computes exactly that trace and
    nothing else; it involves no kernel of this package, so it is checked here against the reference itself."""
    import tamtr_b200
    from tamtr_b200 import modules
    ns = reference_loader.hot_path()
    torch.manual_seed(3)
    ref = ns.TIAGELAN(64, 64, 64, 32, 1, nh=2)
    ours = modules.TIAGELAN(64, 64, 64, 32, 1, nh=2)
    assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    for m in ref.modules():                                     # non-trivial BatchNorm state
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_()
            m.running_var.uniform_(0.5, 2.0)
            m.weight.data.normal_(1.0, 0.2)
            m.bias.data.normal_(0.0, 0.2)
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(2, 64, 12, 10)
    g1, g2 = torch.randn(2, 10, 512), torch.randn(2, 10, 512)
    for mode in ("train", "eval"):
        getattr(ref, mode)()
        getattr(ours, mode)()
        y_ref = ref(x, g1)
        y_ours = ours(x, g2)                                    # a DIFFERENT guide: the output must not care
        assert torch.allclose(y_ours, y_ref, rtol=1e-5, atol=1e-6), mode
        sd_r, sd_o = ref.state_dict(), ours.state_dict()
        for k in sd_r:                                          # includes attn.proj_conv.bn.running_* / num_batches_tracked
            assert torch.allclose(sd_o[k].float(), sd_r[k].float(), rtol=1e-5, atol=1e-6), (mode, k)
    assert int(ref.attn.proj_conv.bn.num_batches_tracked) == 1
    # the same through enable(): the reference's own instance, our forward
    ref2 = ns.TIAGELAN(64, 64, 64, 32, 1, nh=2)
    ref2.load_state_dict(ref.state_dict())
    ref.train(); ref2.train()
    want = ref(x, g1)
    tamtr_b200.enable()
    try:
        got = ref2(x, g2)
    finally:
        tamtr_b200.disable()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ref2.attn.proj_conv.bn.running_mean, ref.attn.proj_conv.bn.running_mean, rtol=1e-5, atol=1e-6)
    assert not any(p.grad is not None for p in ref2.attn.parameters())
