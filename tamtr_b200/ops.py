"""Autograd-aware Python entry points over the C ABI (include/tamtr_b200.h).  CUDA only, no fallback.

`ms_deform_attn` has the signature of the reference's free function
ultralytics/nn/modules/utils.py:42 multi_scale_deformable_attn_pytorch(value, value_spatial_shapes,
sampling_locations, attention_weights) so that tamtr_b200.enable() can rebind that name to it.
"""
import torch

from . import _lib


def _with_device(t):
    return torch.cuda.device(t.device)


class _MSDeformAttnFn(torch.autograd.Function):
    """value [B,Lv,H,Dh] (f32|bf16), loc [B,Lq,H,L,P,2] f32, attn [B,Lq,H,L,P] f32 -> out [B,Lq,H*Dh]."""

    @staticmethod
    def forward(ctx, value, loc, attn, shapes):
        _lib.require_cuda(value, loc, attn)
        value = value.contiguous()
        loc = loc.contiguous().float()
        attn = attn.contiguous().float()
        B, Lv, H, Dh = value.shape
        _, Lq, _, L, P, _ = loc.shape
        sh, nl = _lib.shapes_array(shapes)
        if nl != L:
            raise RuntimeError(f"tamtr_b200: {nl} level shapes given but sampling_locations has {L} levels")
        out = torch.empty(B, Lq, H * Dh, dtype=value.dtype, device=value.device)
        with _with_device(value):
            rc = _lib.lib().tamtr_msda_forward(value.data_ptr(), loc.data_ptr(), attn.data_ptr(), out.data_ptr(),
                                               _lib.dtype_code(value), B, Lv, H, Dh, Lq, L, P, sh,
                                               _lib.stream_ptr(value.device))
        _lib.check(rc, "msda_forward")
        ctx.save_for_backward(value, loc, attn)
        ctx.shapes = [list(map(int, s)) for s in (shapes.tolist() if isinstance(shapes, torch.Tensor) else shapes)]
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        value, loc, attn = ctx.saved_tensors
        grad_out = grad_out.contiguous().to(value.dtype)
        B, Lv, H, Dh = value.shape
        _, Lq, _, L, P, _ = loc.shape
        sh, _ = _lib.shapes_array(ctx.shapes)
        grad_value = torch.empty_like(value)          # zeroed by the C call
        grad_loc = torch.empty_like(loc)
        grad_attn = torch.empty_like(attn)
        # grad_value is accumulated with vector atomics -> run-to-run bit differences, exactly like the
        # reference's grid_sampler_2d_backward; honour torch.use_deterministic_algorithms the way ATen does.
        if torch.are_deterministic_algorithms_enabled():
            if torch.is_deterministic_algorithms_warn_only_enabled():
                import warnings
                warnings.warn("tamtr_b200 ms_deform_attn backward uses atomicAdd and is not deterministic")
            else:
                raise RuntimeError("tamtr_b200 ms_deform_attn backward does not have a deterministic implementation")
        with _with_device(value):
            rc = _lib.lib().tamtr_msda_backward(grad_out.data_ptr(), value.data_ptr(), loc.data_ptr(),
                                                attn.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
                                                grad_attn.data_ptr(), _lib.dtype_code(value), B, Lv, H, Dh, Lq, L, P,
                                                sh, _lib.stream_ptr(value.device))
        _lib.check(rc, "msda_backward")
        return grad_value, grad_loc, grad_attn, None


def ms_deform_attn(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Drop-in for multi_scale_deformable_attn_pytorch (utils.py:42).  fp16 values are computed in fp32."""
    _lib.require_cuda(value, sampling_locations, attention_weights)
    if value.dtype == torch.float16 or value.dtype == torch.float64:
        out = _MSDeformAttnFn.apply(value.float(), sampling_locations, attention_weights, value_spatial_shapes)
        return out.to(value.dtype)
    return _MSDeformAttnFn.apply(value, sampling_locations, attention_weights, value_spatial_shapes)


def ms_deform_attn_corners(sampling_locations, value_spatial_shapes):
    """Parity export of the kernel's index math: (x0, y0) int32 [B,Lq,H,L,P] and in-bounds flags uint8 [...,4]."""
    _lib.require_cuda(sampling_locations)
    loc = sampling_locations.contiguous().float()
    B, Lq, H, L, P, _ = loc.shape
    sh, _ = _lib.shapes_array(value_spatial_shapes)
    x0 = torch.empty(B, Lq, H, L, P, dtype=torch.int32, device=loc.device)
    y0 = torch.empty_like(x0)
    inb = torch.empty(B, Lq, H, L, P, 4, dtype=torch.uint8, device=loc.device)
    with _with_device(loc):
        rc = _lib.lib().tamtr_msda_corners(loc.data_ptr(), x0.data_ptr(), y0.data_ptr(), inb.data_ptr(),
                                           B, Lq, H, L, P, sh, _lib.stream_ptr(loc.device))
    _lib.check(rc, "msda_corners")
    return x0, y0, inb
