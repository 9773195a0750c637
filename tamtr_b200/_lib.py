"""ctypes binding of the C ABI in include/tamtr_b200.h (libtamtr_b200.so, built in-tree for sm_100a).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.  CPU tensors are
rejected with a message containing 'Not implemented on the CPU' / 'is_cuda', which is what the reference's
DetectionModel.__init__ dry-run looks for before moving the model to CUDA (ultralytics/nn/tasks.py:256-264).
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libtamtr_b200.so")
CSRC = os.path.join(_HERE, "csrc")

F32, BF16 = 0, 1
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}

_lib = None

_vp, _i, _fp, _l = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_long
_SIGNATURES = {
    "tamtr_abi_version": (ctypes.c_int, []),
    "tamtr_last_error": (ctypes.c_char_p, []),
    "tamtr_launch_count": (ctypes.c_ulonglong, []),
    "tamtr_memset_zero": (ctypes.c_int, [_vp, ctypes.c_ulonglong, _vp]),
    "tamtr_zero_fill_background": (ctypes.c_int, [_vp, ctypes.c_ulonglong, _i, _vp]),
    "tamtr_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "tamtr_profile_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                                          ctypes.POINTER(ctypes.c_ulonglong)]),
    "tamtr_kernel_name": (ctypes.c_char_p, [ctypes.c_int]),
    "tamtr_msda_forward": (ctypes.c_int, [_vp, _fp, _fp, _vp, _i] + [_i] * 7 + [_vp, _i, _vp]),
    "tamtr_msda_backward": (ctypes.c_int, [_vp, _vp, _fp, _fp, _vp, _fp, _fp, _i] + [_i] * 7 + [_vp, _i, _i, _fp, _i, _vp]),
    "tamtr_msda_corners": (ctypes.c_int, [_fp, _vp, _vp, _vp] + [_i] * 5 + [_vp, _vp]),
    "tamtr_msda_forward_ragged": (ctypes.c_int, [_vp, _fp, _fp, _vp, _i] + [_i] * 6 + [_vp, _vp, _i, _vp]),
    "tamtr_msda_backward_ragged": (ctypes.c_int, [_vp, _vp, _fp, _fp, _vp, _fp, _fp, _i] + [_i] * 6
                                   + [_vp, _vp, _i, _i, _fp, _i, _vp]),
    "tamtr_msda_corners_ragged": (ctypes.c_int, [_fp, _vp, _vp, _vp] + [_i] * 4 + [_vp, _vp, _vp]),
    "tamtr_locw_forward": (ctypes.c_int, [_fp] * 5 + [_i] * 6 + [_vp, _vp]),
    "tamtr_locw_backward": (ctypes.c_int, [_fp] * 8 + [_i] * 6 + [_vp, _vp]),
    "tamtr_box_refine_forward": (ctypes.c_int, [_fp, _fp, _fp, _i, ctypes.c_float, _vp]),
    "tamtr_box_refine_backward": (ctypes.c_int, [_fp, _fp, _fp, _fp, _fp, _i, ctypes.c_float, _vp]),
    "tamtr_contrastive_forward": (ctypes.c_int, [_vp, _fp, _fp, _fp, _fp] + [_i] * 5 + [_vp]),
    "tamtr_contrastive_backward": (ctypes.c_int, [_fp, _vp, _fp, _fp, _vp, _fp] + [_i] * 5 + [_vp]),
    "tamtr_max_sigmoid_forward": (ctypes.c_int, [_vp, _fp, _fp, _fp, _vp] + [_i] * 6 + [_vp]),
    "tamtr_col_reduce2_ctas": (ctypes.c_int, [_i, _i]),
    "tamtr_col_reduce2": (ctypes.c_int, [_vp, _vp, _fp] + [_i] * 6 + [_vp]),
    "tamtr_affine_rows": (ctypes.c_int, [_vp, _vp, _vp, _fp, _fp, _fp] + [_i] * 5 + [_vp, _vp]),
    "tamtr_rank_tokens": (ctypes.c_int, [_vp, _fp, _fp, _vp, _fp, _fp, _fp, _fp] + [_i] * 6 + [ctypes.c_float, _vp]),
    "tamtr_max_sigmoid_tc_forward": (ctypes.c_int, [_vp, _fp, _fp, _fp, _vp] + [_i] * 5 + [_vp]),
    "tamtr_gate_conv3x3_tc_forward": (ctypes.c_int, [_vp, _vp, _fp, _fp, _fp, _vp] + [_i] * 6 + [_vp]),
    "tamtr_nchw_to_nhwc": (ctypes.c_int, [_vp, _vp] + [_i] * 4 + [_vp]),
    "tamtr_topk_rows_supported": (ctypes.c_int, [_i, _i]),
    "tamtr_topk_rows": (ctypes.c_int, [_fp, _vp, _fp, _i, _i, _i, _vp]),
    "tamtr_add_layernorm_forward": (ctypes.c_int, [_vp, _i, _vp, _i, _fp, _fp, _vp, _i, _fp, _fp, _fp, _i, _i,
                                                   ctypes.c_float, _vp]),
    "tamtr_add_layernorm_backward": (ctypes.c_int, [_vp, _i, _fp, _fp, _fp, _fp, _vp, _i, _vp, _i, _fp, _i, _i, _vp]),
    "tamtr_add_layernorm_forward_sides": (ctypes.c_int, [_vp, _i, _vp, _i, _fp, _fp, _vp, _i, _fp, _fp, _fp, _vp, _i, _vp, _vp,
                                                         _i, _i, ctypes.c_float, _vp]),
    "tamtr_add_layernorm_backward_sides": (ctypes.c_int, [_vp, _i, _vp, _vp, _fp, _fp, _fp, _fp, _vp, _i, _vp, _i, _fp, _i, _i,
                                                          _vp]),
    "tamtr_pos_cast": (ctypes.c_int, [_vp, _i, _vp, _i, _vp, _vp, _l, _vp]),
    "tamtr_grad_sum3": (ctypes.c_int, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _l, _vp]),
    "tamtr_add_layernorm_backward_res": (ctypes.c_int, [_vp, _i, _fp, _fp, _fp, _fp, _fp, _vp, _i, _vp, _i, _fp, _i, _i, _vp]),
    "tamtr_self_attention_supported": (ctypes.c_int, [_i] * 3),
    "tamtr_self_attention_padded_len": (ctypes.c_int, [_i]),
    "tamtr_self_attention_mask_words": (ctypes.c_int, [_i]),
    "tamtr_self_attention_pack_mask": (ctypes.c_int, [_vp, _vp, _i, _vp]),
    "tamtr_self_attention_forward": (ctypes.c_int, [_vp, _vp, _vp, _vp, _fp] + [_i] * 4 + [_vp]),
    "tamtr_self_attention_backward": (ctypes.c_int, [_vp] * 5 + [_fp, _vp, _vp, _vp] + [_i] * 4 + [_vp]),
    "tamtr_bn_forward_coeffs": (ctypes.c_int, [_fp, _i, ctypes.c_double, _fp, _fp, ctypes.c_double, _i, ctypes.c_double,
                                               _fp, _fp, _fp, _fp, _vp, _vp, _i, _vp]),
    "tamtr_bn_backward_coeffs": (ctypes.c_int, [_fp, _i, ctypes.c_double, _fp, _vp, _vp, _i, _fp, _fp, _fp, _fp, _fp,
                                                _i, _vp]),
    "tamtr_col_sum": (ctypes.c_int, [_vp, _fp, _i, _i, _i, _vp]),
    "tamtr_linear_sum_assignment": (ctypes.c_int, [_fp, _vp, _vp, _vp, _vp] + [_i] * 6 + [ctypes.c_longlong, _vp]),
    "tamtr_selective_scan_segments": (ctypes.c_int, [_i]),
    "tamtr_selective_scan_forward": (ctypes.c_int, [_vp, _vp, _i] + [_fp] * 7 + [_i] * 5 + [_vp]),
    "tamtr_selective_scan_backward": (ctypes.c_int, [_vp, _vp, _i] + [_fp] * 7 + [_vp, _vp] + [_fp] * 5 + [_i] * 5 + [_vp]),
    "tamtr_locw_tc_supported": (ctypes.c_int, [_i] * 7),
    "tamtr_locw_tc_forward": (ctypes.c_int, [_vp] * 5 + [_i] + [_fp] * 4 + [_i] * 7 + [_vp]),
    "tamtr_dwconv3x3_silu_forward": (ctypes.c_int, [_vp, _fp, _fp, _vp] + [_i] * 5 + [_vp]),
    "tamtr_dwconv3x3_silu_backward": (ctypes.c_int, [_vp, _vp, _fp, _fp, _vp, _fp, _fp] + [_i] * 5 + [_vp]),
    "tamtr_selective_scan_chunks": (ctypes.c_int, [_i] * 3),
    "tamtr_selective_scan_forward_chunked": (ctypes.c_int, [_vp, _vp, _i] + [_fp] * 7 + [_i] * 6 + [_vp]),
    "tamtr_colnorm_gate_forward": (ctypes.c_int, [_fp, _vp, _i, _fp, _fp, _vp, _i, _fp, _fp, _i, _i, _i, ctypes.c_float, _vp]),
    "tamtr_colnorm_gate_backward": (ctypes.c_int, [_vp, _i, _fp, _vp, _i, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _i, _i, _i, _vp]),
    "tamtr_cross_scan": (ctypes.c_int, [_vp, _vp] + [_i] * 5 + [_vp]),
    "tamtr_cross_merge": (ctypes.c_int, [_vp, _vp] + [_i] * 5 + [_vp]),
    "tamtr_cdn_group": (ctypes.c_int, [_fp, _vp, _vp, _fp, _vp, _fp, _fp, _vp] + [_i] * 6 + [ctypes.c_float] * 2 + [_vp]),
    "tamtr_match_cost": (ctypes.c_int, [_fp, _fp, _fp, _vp, _fp] + [_i] * 5 + [ctypes.c_float] * 5 + [_vp]),
    "tamtr_linear_sum_assignment_padded": (ctypes.c_int, [_fp, _vp, _vp] + [_i] * 4 + [_vp]),
    "tamtr_detection_loss": (ctypes.c_int, [_fp, _fp, _fp, _vp, _vp, _vp, _fp, _fp, _fp, _fp, _fp] + [_i] * 8
                             + [ctypes.c_float] * 3 + [_vp]),
    "tamtr_optim_partials": (ctypes.c_int, [ctypes.c_longlong]),
    "tamtr_adamw_flat": (ctypes.c_int, [_fp, _fp, _fp, _fp, ctypes.c_longlong, _vp, _fp, _fp]
                         + [ctypes.c_float] * 6 + [_vp]),
    "tamtr_tok_project_supported": (ctypes.c_int, [_i] * 6),
    "tamtr_tok_project": (ctypes.c_int, [_vp, _vp, _fp, _vp, _vp, _l, _l, _vp, _l, _l, _fp, _l, _l] + [_i] * 6 + [_vp]),
    "tamtr_tok_project_rank": (ctypes.c_int, [_vp, _vp, _fp, _vp, _vp, _l, _l, _fp, _l, _vp, _fp, _i, ctypes.c_float] + [_i] * 6 + [_vp]),
    "tamtr_fold_rank_consts": (ctypes.c_int, [_vp] * 4 + [_fp] * 4 + [_i] * 5 + [_vp]),
    "tamtr_fold_stats": (ctypes.c_int, [_i] + [_vp] * 7 + [_vp]),
    "tamtr_fold_bn": (ctypes.c_int, [_i, _i] + [_vp] * 12 + [_i, _i] + [_fp] * 3 + [_vp]),
    "tamtr_fold_pack": (ctypes.c_int, [_i, _vp, _fp, _fp, _fp, _vp, _fp, _i, _i, _vp]),
    "tamtr_fold_unpack": (ctypes.c_int, [_i] + [_vp] * 4 + [_fp] * 3 + [_i, _vp]),
    "tamtr_fold_bn_bwd": (ctypes.c_int, [_i, _i] + [_vp] * 5 + [_fp] * 3 + [_i] + [_vp] * 3 + [_fp, _vp]),
    "tamtr_fold_gather": (ctypes.c_int, [_i, _i] + [_vp] * 4 + [_vp, _fp, _i, _vp]),
    "tamtr_tok_reduce_supported": (ctypes.c_int, [_i] * 5),
    "tamtr_tok_reduce_splits": (ctypes.c_int, [_i] * 5),
    "tamtr_tok_reduce": (ctypes.c_int, [_vp, _l, _l, _i, _vp, _fp, _fp] + [_i] * 4 + [_vp]),
    "tamtr_max_sigmoid_backward": (ctypes.c_int, [_fp, _fp, _vp, _vp, _fp, _vp, _fp, _fp] + [_i] * 6 + [_vp]),
}


def build(verbose=False):
    """Compile every CUDA source for sm_100a into tamtr_b200/lib/libtamtr_b200.so (nvcc cross-compiles on CPU)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("tamtr_b200: nvcc build failed\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def exported_symbols():
    return list(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"tamtr_b200: {LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        if handle.tamtr_abi_version() != 2:
            raise RuntimeError("tamtr_b200: ABI version mismatch between _lib.py and libtamtr_b200.so")
        _lib = handle
    return _lib


def launch_count():
    return int(lib().tamtr_launch_count())


def zeros_like_fast(t):
    """torch.zeros_like through a memset node instead of a fill kernel."""
    out = torch.empty(t.shape, dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        check(lib().tamtr_memset_zero(out.data_ptr(), out.numel() * out.element_size(), stream_ptr(t.device)),
              "memset_zero")
    return out


def profile_enable(on=True):
    """Bracket every kernel launch of the library with CUDA events on its stream (clears previous records)."""
    lib().tamtr_profile_enable(1 if on else 0)


def profile_read():
    """-> {kernel_name: (total_ms, launches)} for kernels launched since profile_enable(True)."""
    out = {}
    for kid in range(64):
        name = lib().tamtr_kernel_name(kid).decode()
        if not name:
            break
        ms, n = ctypes.c_double(0), ctypes.c_ulonglong(0)
        lib().tamtr_profile_read(kid, ctypes.byref(ms), ctypes.byref(n))
        if n.value:
            out[name] = (ms.value, int(n.value))
    return out


def check(rc, what):
    if rc != 0:
        msg = lib().tamtr_last_error().decode(errors="replace")
        raise RuntimeError(f"tamtr_b200: {what} failed (code {rc}): {msg}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("tamtr_b200: Not implemented on the CPU (expected is_cuda tensors); "
                               "this path has hand-written sm_100a kernels only")


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"tamtr_b200: dtype {t.dtype} not supported (float32 / bfloat16)") from None


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def shapes_array(shapes):
    """value_spatial_shapes arrives as a python list of [h, w] (head.py:1215) or a tensor -> host int32[L][2]."""
    if isinstance(shapes, torch.Tensor):
        shapes = shapes.tolist()
    flat = []
    for h, w in shapes:
        flat += [int(h), int(w)]
    return (ctypes.c_int32 * len(flat))(*flat), len(flat) // 2
