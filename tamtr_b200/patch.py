"""Install / remove the CUDA path inside an imported reference `ultralytics` package.

    import tamtr_b200; tamtr_b200.enable()          # before or after the model is built; one line in trainTAMTR.py

Classes are never replaced -- only attributes of the existing classes are (re)bound -- so `isinstance(m, TIAGELAN)`
(nn/tasks.py:662), pickled checkpoints (`ultralytics.nn.modules.transformer.MSDeformAttn` in TAM_TR.pt), `copy.deepcopy`
(EMA, _get_clones) and state_dict keys keep working unchanged.  What enable() installs is the
WHOLE fast path, not only its leaves:

  core ops      multi_scale_deformable_attn_pytorch / _cls / _box in nn/modules/utils.py AND the names imported into
                nn/modules/transformer.py (utils.py:42,92,143; transformer.py:12)              -> ops.ms_deform_attn*
  attention     MSDeformAttn / MSDeformAttncls / MSDeformAttnbox .forward (transformer.py:252,347,445)
  layers        DeformableTransformerDecoderLayer / DecouplingDeformableTransformerDecoderLayer .forward (:539,:621)
  decoders      DeformableTransformerDecoder / TextDeformableTransformerDecoder .forward (:681,:850): value projections of
                all layers as one GEMM + shared gradient arena, fused box refinement
  heads         RTDETRDecoder / ManbaWorldDecoder .forward, ._get_encoder_input, ._get_decoder_input, ._generate_anchors
                (nn/modules/head.py:280-435, 1130-1264): token-major input projection with fused BatchNorm, ranking kernel,
                sparse query selection, cached anchors, device-side denoising-group materialisation
  score head    ContrastiveHeadMLP.forward (nn/modules/block.py:534)
  BTA-PAN       MaxSigmoidAttnBlock.forward (extra_modules/block.py:208 and the copy in modules/block.py),
                TIAGELAN.forward / forward_split (extra_modules/block.py:182-192: the block's output is discarded there)
  VSSBlocks     VSSBlock.forward, SS2D.forwardv2 (extra_modules/VManba/vmamba.py:1236-1257, 1019-1038) on the selective-scan
                kernels; the missing `selective_scan_cuda_core` extension (csms6s.py:252-270) is supplied as a shim over
                the same kernels, so the reference's own SelectiveScanCore also runs
  loss          HungarianMatcher.forward (models/utils/ops.py:48-121) -> device-side assignment; DETRLoss.forward /
                RTDETRDetectionLoss.forward (models/utils/loss.py:328-443) -> all layers in one set of kernels

CPU tensors raise "Not implemented on the CPU ... is_cuda" (no fallback); nn/tasks.py:256-264 reacts to exactly that
message by moving the model to CUDA for its construction-time dry run.  Export/tracing (engine/exporter.py) is
reference-path only: call disable() first.
"""
import importlib
import inspect

from . import head, loss, modules, ops, vss

_saved = []
_MISSING = object()


def _rebind(obj, name, new):
    _saved.append((obj, name, obj.__dict__.get(name, _MISSING) if isinstance(obj, type) else getattr(obj, name, _MISSING)))
    setattr(obj, name, new)


def install(ref_cls, ours_cls, names, rebind=None):
    """Bind the attributes `names` of our mirror class onto the reference's class (functions stay plain functions,
    staticmethods stay staticmethods).  `rebind`: how to set one attribute (enable() records the old value; tests that
    build reference-shaped stand-ins pass plain setattr)."""
    rebind = rebind or (lambda o, n, v: setattr(o, n, v))
    for name in names:
        rebind(ref_cls, name, inspect.getattr_static(ours_cls, name))


HEAD_ATTRS = ("forward", "_get_encoder_input", "_get_decoder_input", "_generate_anchors", "_encode", "_cdn", "_anchors",
              "_rank_tokens", "_fusable_input_proj", "_finish", "plan_cdn", "fused_input_proj", "sparse_query_selection",
              "__getstate__", "_valid_u8", "_fold_attns", "folded_projection")
DECODER_ATTRS = ("forward", "_run", "_project_values", "batched_value_projection")
MATCHER_ATTRS = ("forward", "match_layers", "match_padded")
LOSS_ATTRS = ("forward", "_get_loss_layers")


def enabled():
    return bool(_saved)


def _optional(package, name):
    try:
        return importlib.import_module(package + name)
    except Exception:       # a stripped-down copy of the package (the test loader imports only the hot path)
        return None


def enable(package="ultralytics"):
    if _saved:
        return
    T = importlib.import_module(package + ".nn.modules.transformer")
    U = importlib.import_module(package + ".nn.modules.utils")
    B = importlib.import_module(package + ".nn.modules.block")
    _rebind(U, "multi_scale_deformable_attn_pytorch", ops.ms_deform_attn)
    _rebind(T, "multi_scale_deformable_attn_pytorch", ops.ms_deform_attn)
    _rebind(T.MSDeformAttn, "forward", modules.MSDeformAttn.forward)
    for name, fn in (("multi_scale_deformable_attn_pytorch_cls", ops.ms_deform_attn_cls),
                     ("multi_scale_deformable_attn_pytorch_box", ops.ms_deform_attn_box)):
        for mod in (U, T):
            if hasattr(mod, name):
                _rebind(mod, name, fn)
    for name in ("MSDeformAttncls", "MSDeformAttnbox", "DecouplingDeformableTransformerDecoderLayer"):
        if hasattr(T, name):
            _rebind(getattr(T, name), "forward", getattr(modules, name).forward)
    _rebind(T.DeformableTransformerDecoderLayer, "forward", modules.DeformableTransformerDecoderLayer.forward)
    for name in ("DeformableTransformerDecoder", "TextDeformableTransformerDecoder"):
        if hasattr(T, name):
            install(getattr(T, name), getattr(modules, name), DECODER_ATTRS, _rebind)
    if hasattr(B, "ContrastiveHeadMLP"):
        _rebind(B.ContrastiveHeadMLP, "forward", modules.ContrastiveHeadMLP.forward)
    if hasattr(B, "MaxSigmoidAttnBlock"):
        _rebind(B.MaxSigmoidAttnBlock, "forward", modules.MaxSigmoidAttnBlock.forward)
    E = _optional(package, ".nn.extra_modules.block")
    if E is not None:
        _rebind(E.MaxSigmoidAttnBlock, "forward", modules.MaxSigmoidAttnBlock.forward)
        if hasattr(E, "TIAGELAN"):
            _rebind(E.TIAGELAN, "forward", modules.TIAGELAN.forward)
            _rebind(E.TIAGELAN, "forward_split", modules.TIAGELAN.forward)
    Hd = _optional(package, ".nn.modules.head")
    if Hd is not None:
        for name in ("RTDETRDecoder", "ManbaWorldDecoder"):
            if hasattr(Hd, name):
                install(getattr(Hd, name), getattr(head, name), HEAD_ATTRS, _rebind)
    V = _optional(package, ".nn.extra_modules.VManba.vmamba")
    if V is not None:
        if hasattr(V, "VSSBlock"):
            _rebind(V.VSSBlock, "forward", vss.vssblock_forward_on(V.VSSBlock.forward))
        if hasattr(V, "SS2D") and hasattr(V.SS2D, "forwardv2"):
            _rebind(V.SS2D, "forwardv2", vss.ss2d_forward_on(V.SS2D.forwardv2))
    C = _optional(package, ".nn.extra_modules.VManba.csms6s")
    if C is not None and not hasattr(C, "selective_scan_cuda_core"):
        _rebind(C, "selective_scan_cuda_core", vss.ScanExtensionShim)
    O = _optional(package, ".models.utils.ops")
    if O is not None and hasattr(O, "HungarianMatcher"):
        install(O.HungarianMatcher, loss.HungarianMatcher, MATCHER_ATTRS, _rebind)
    Ls = _optional(package, ".models.utils.loss")
    if Ls is not None and hasattr(Ls, "DETRLoss"):
        install(Ls.DETRLoss, loss.DETRLoss, LOSS_ATTRS, _rebind)
        if hasattr(Ls, "RTDETRDetectionLoss"):
            _rebind(Ls.RTDETRDetectionLoss, "forward", loss.RTDETRDetectionLoss.forward)


def disable():
    while _saved:
        obj, name, old = _saved.pop()
        if old is _MISSING:
            delattr(obj, name)
        else:
            setattr(obj, name, old)
