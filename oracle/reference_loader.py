"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the UNMODIFIED reference (/root/reference, a fork of Ultralytics 8.0.201) in THIS container so that
`oracle/make_goldens.py` can generate golden vectors from it.  /root/reference does not exist on the GPU box, so
nothing under tests/ -m gpu, smoke() or bench.py may call this module at run time.

The reference cannot be imported as shipped: `ultralytics/utils/__init__.py:19` needs matplotlib and
`ultralytics/nn/extra_modules/block.py:17` needs timm; fvcore, thop, seaborn, clip and dill are also absent.
None of them are on the hot path, so they are replaced by empty stand-ins before import.
"""
import importlib
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("TAMTR_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Attribute sink: any attribute/call on a stubbed third-party package yields another sink."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave as a package so that `import a.b` works
    mod.__dict__.update(attrs)

    def _module_getattr(attr):  # PEP 562
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Anything()

    mod.__getattr__ = _module_getattr
    sys.modules[name] = mod
    return mod


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics"))


_loaded = None


def load():
    """Return the imported reference `ultralytics` package (with out-of-path third-party deps stubbed)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT} (it only exists in the build container)")
    import torch
    import torch.nn as nn

    tmp = tempfile.mkdtemp(prefix="tamtr_ref_home_")
    os.environ.setdefault("YOLO_CONFIG_DIR", tmp)
    os.environ.setdefault("HOME", tmp)

    class DropPath(nn.Module):  # timm.layers.DropPath, eval/zero-prob behaviour only
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            return x * mask / keep

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.colors",
                 "seaborn", "thop", "fvcore", "fvcore.nn", "clip", "dill"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    for name in ("timm", "timm.layers", "timm.models", "timm.models.layers"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name, DropPath=DropPath, trunc_normal_=nn.init.trunc_normal_,
                      to_2tuple=lambda v: v if isinstance(v, tuple) else (v, v))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    torch.set_grad_enabled(True)
    _loaded = importlib.import_module("ultralytics")
    return _loaded


def hot_path():
    """Namespace with the reference callables on the hot path."""
    load()
    ns = types.SimpleNamespace()
    u = importlib.import_module("ultralytics.nn.modules.utils")
    t = importlib.import_module("ultralytics.nn.modules.transformer")
    b = importlib.import_module("ultralytics.nn.modules.block")
    h = importlib.import_module("ultralytics.nn.modules.head")
    e = importlib.import_module("ultralytics.nn.extra_modules.block")
    o = importlib.import_module("ultralytics.models.utils.ops")
    ns.utils, ns.transformer, ns.block, ns.head, ns.extra_block, ns.ops = u, t, b, h, e, o
    ns.multi_scale_deformable_attn_pytorch = u.multi_scale_deformable_attn_pytorch
    ns.inverse_sigmoid = u.inverse_sigmoid
    ns.MSDeformAttn = t.MSDeformAttn
    ns.MLP = t.MLP
    ns.DeformableTransformerDecoderLayer = t.DeformableTransformerDecoderLayer
    ns.DeformableTransformerDecoder = t.DeformableTransformerDecoder
    ns.TextDeformableTransformerDecoder = t.TextDeformableTransformerDecoder
    ns.ContrastiveHeadMLP = b.ContrastiveHeadMLP
    ns.MaxSigmoidAttnBlock = e.MaxSigmoidAttnBlock
    ns.TIAGELAN = e.TIAGELAN
    ns.RTDETRDecoder = h.RTDETRDecoder
    ns.ManbaWorldDecoder = h.ManbaWorldDecoder
    ns.get_cdn_group = o.get_cdn_group
    return ns
