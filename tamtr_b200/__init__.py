"""tamtr_b200 -- B200-native (sm_100a) drop-in for the TAM-TR detection-head hot path.

Public surface
  ops.ms_deform_attn(value, value_spatial_shapes, sampling_locations, attention_weights)
      = ultralytics/nn/modules/utils.py:42 multi_scale_deformable_attn_pytorch
  enable() / disable()   install / remove the CUDA path inside an imported reference `ultralytics` package
  build()                compile tamtr_b200/lib/libtamtr_b200.so (C ABI in include/tamtr_b200.h)

The product path never imports `oracle/` and has no CPU fallback.
"""
from . import _lib, ops  # noqa: F401
from ._lib import build, launch_count  # noqa: F401
from .ops import ms_deform_attn  # noqa: F401
from .patch import disable, enable, enabled  # noqa: F401

__version__ = "0.1.0"
