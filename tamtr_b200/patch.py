"""Install / remove the CUDA path inside an imported reference `ultralytics` package.

    import tamtr_b200; tamtr_b200.enable()          # before or after the model is built; one line in trainTAMTR.py

Classes are never replaced -- only methods are rebound -- so `isinstance(m, TIAGELAN)` (nn/tasks.py:662), pickled
checkpoints (`ultralytics.nn.modules.transformer.MSDeformAttn` in TAM_TR.pt), `copy.deepcopy` (EMA, _get_clones) and
state_dict keys keep working unchanged.  What is rebound:

  ultralytics.nn.modules.utils.multi_scale_deformable_attn_pytorch        (utils.py:42)       -> ops.ms_deform_attn
  ultralytics.nn.modules.transformer.multi_scale_deformable_attn_pytorch  (name imported at transformer.py:12)
  MSDeformAttn.forward                        (transformer.py:252)  -> fused projection epilogue + sampler
  multi_scale_deformable_attn_pytorch_cls / _box (utils.py:92,143)  -> ops.ms_deform_attn_cls / _box (ragged sampler)
  MSDeformAttncls / MSDeformAttnbox .forward  (transformer.py:347,445), DecouplingDeformableTransformerDecoderLayer.forward (:621)
  DeformableTransformerDecoderLayer.forward   (transformer.py:539)  -> same math, need_weights=False self-attention
  ContrastiveHeadMLP.forward                  (block.py:534)        -> fused contrastive head
  MaxSigmoidAttnBlock.forward                 (extra_modules/block.py:208 and its copy in modules/block.py)

CPU tensors raise "Not implemented on the CPU ... is_cuda" (no fallback); nn/tasks.py:256-264 reacts to exactly that
message by moving the model to CUDA for its construction-time dry run.  Export/tracing (engine/exporter.py) is
reference-path only: call disable() first.
"""
import importlib

from . import modules, ops

_saved = []


def _rebind(obj, name, new):
    _saved.append((obj, name, getattr(obj, name)))
    setattr(obj, name, new)


def enabled():
    return bool(_saved)


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
    if hasattr(B, "ContrastiveHeadMLP"):
        _rebind(B.ContrastiveHeadMLP, "forward", modules.ContrastiveHeadMLP.forward)
    if hasattr(B, "MaxSigmoidAttnBlock"):
        _rebind(B.MaxSigmoidAttnBlock, "forward", modules.MaxSigmoidAttnBlock.forward)
    try:
        E = importlib.import_module(package + ".nn.extra_modules.block")
        _rebind(E.MaxSigmoidAttnBlock, "forward", modules.MaxSigmoidAttnBlock.forward)
    except ImportError:
        pass


def disable():
    while _saved:
        obj, name, old = _saved.pop()
        setattr(obj, name, old)
