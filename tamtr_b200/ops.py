"""Autograd-aware Python entry points over the C ABI (include/tamtr_b200.h).  CUDA only, no fallback.

`ms_deform_attn` has the signature of the reference's free function
ultralytics/nn/modules/utils.py:42 multi_scale_deformable_attn_pytorch(value, value_spatial_shapes,
sampling_locations, attention_weights) so that tamtr_b200.enable() can rebind that name to it.
"""
import ctypes
import os

import torch

from . import _lib


def _with_device(t):
    return torch.cuda.device(t.device)


def _value_layout(value):
    """value [B, Lv, H, Dh] either contiguous or a column slice of a wider [B, Lv, C_total] buffer (batched value
    projection).  Returns (tensor to pass, token stride in elements)."""
    B, Lv, H, Dh = value.shape
    st = value.stride()
    if st[3] == 1 and st[2] == Dh and st[1] >= H * Dh and (st[0] == Lv * st[1] or B == 1) \
            and (st[1] * value.element_size()) % 16 == 0 and (value.storage_offset() * value.element_size()) % 16 == 0:
        return value, st[1]
    value = value.contiguous()
    return value, H * Dh


def _env_arena_dtype():
    v = os.environ.get("TAMTR_GRAD_ARENA", "").lower()
    return torch.float32 if v in ("fp32", "f32", "float32") else None


ARENA_PREFILL = os.environ.get("TAMTR_ARENA_PREFILL", "1") != "0"
# CTAs of the forked fill kernel: 0 = one per SM (measured at S-yaml B=16: 64 CTAs -0.22 ms, 148 CTAs -0.23 ms per step,
# 32 CTAs +0.07 ms: the window up to the first sampler backward is ~0.7 ms); "memset" = a cudaMemsetAsync node on the side
# stream (no gain: +0.03 ms -- a full-grid memset beside the decoder hides nothing)
ARENA_FILL_CTAS = os.environ.get("TAMTR_ARENA_FILL_CTAS", "0")
ARENA_FILL_CTAS = ARENA_FILL_CTAS if ARENA_FILL_CTAS == "memset" else int(ARENA_FILL_CTAS)
# fill kernel: "bulk" = bulk stores from a zeroed 32 KB shared tile; "regs" = 16-byte stores from registers, no shared memory
# (measured equal: both give 4.21 ms on most processes and 4.02 ms on some)
ARENA_FILL_KERNEL = os.environ.get("TAMTR_ARENA_FILL_KERNEL", "bulk")
ARENA_DEFER = os.environ.get("TAMTR_ARENA_DEFER", "1") != "0"     # fork the fill behind the query selection's top-k
_SIDE_STREAMS = {}


def _side_stream(device):
    """One side stream per device for work forked off the step (the gradient arena's zero fill)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(idx)
    if st is None:
        st = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx)
    return st


GRAD_ARENA_DTYPE = _env_arena_dtype()     # default gradient dtype of ValueArena (None = value dtype); TAMTR_GRAD_ARENA=fp32


class ValueArena:
    """Shared gradient state for the value tensors of all decoder layers.

    The decoder projects `feats` once for every layer (one [d, n_layers*d] GEMM); each layer's sampler reads its
    column slice.  In the backward every sampler accumulates into ITS slice of one zero-initialised
    [B, Lv, n_layers*d] buffer (memset node), which then feeds a single dgrad and a single wgrad GEMM -- no per-layer
    dense gradient tensors and no gradient-accumulation passes over [B, Lv, d].  The bias gradient (column sums of
    that buffer) is assembled from the samplers' per-(query, head) tap-weight sums instead of re-reading it."""

    def __init__(self, grad_dtype=None):
        self.buf = None
        self._ready = None        # event of a zero fill forked to the side stream (prefill), pending a join
        self.defer = False        # prefill() only notes the request; start_prefill() forks the fill (see there)
        self._pending = None
        self.base = None
        self.bias_grad = {}       # column offset of a layer's slice -> [d] fp32
        self.written = set()      # column offsets whose sampler backward has run in this backward pass
        # dtype of the gradient buffer: None = the value dtype (bf16 values -> bf16x8 vector reductions);
        # torch.float32 = fp32 accumulation of the scattered gradient (f32x4 reductions, one cast pass before the GEMMs)
        self.grad_dtype = grad_dtype if grad_dtype is not None else GRAD_ARENA_DTYPE

    def prefill(self, shape, dtype, device):
        """Zero-fill the gradient buffer AHEAD of the backward, on a side stream forked from the caller's stream here
        (call it right after the value projection was launched) and joined by the first consumer (`grad_buffer` /
        `take`).  The 1.65 GB memset node of the S-yaml step is pure HBM write time (0.26 ms); on the main stream it sits
        in front of the first sampler backward with nothing beside it, while the decoder forward between the projection
        and that point is latency-bound small kernels that leave the memory system idle.  In a captured step the fork
        and the join become parallel branches of the graph.  Only call it when a backward will follow: a capture that
        ends with the branch unjoined is an error (TAMTR_ARENA_PREFILL=0 keeps the memset in the backward)."""
        if not ARENA_PREFILL or self.buf is not None or device.type != "cuda":
            return
        if self.defer and ARENA_DEFER:
            self._pending = (tuple(shape), dtype, device)
            return
        gdt = self.grad_dtype if self.grad_dtype is not None else dtype
        buf = torch.empty(shape, dtype=gdt, device=device)                   # allocated on (and owned by) the main stream
        main, side = torch.cuda.current_stream(device), _side_stream(device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            nbytes = buf.numel() * buf.element_size()
            if ARENA_FILL_CTAS != "memset":     # small co-resident CTAs beside the decoder's kernels, not a full-grid memset
                n_ctas = ARENA_FILL_CTAS
                if ARENA_FILL_KERNEL == "regs" and n_ctas >= 0:
                    n_ctas = -(n_ctas or torch.cuda.get_device_properties(device).multi_processor_count)
                _lib.check(_lib.lib().tamtr_zero_fill_background(buf.data_ptr(), nbytes, n_ctas,
                                                                 _lib.stream_ptr(device)), "zero_fill_background")
            else:
                _lib.check(_lib.lib().tamtr_memset_zero(buf.data_ptr(), nbytes, _lib.stream_ptr(device)), "memset_zero")
            self._ready = torch.cuda.Event()
            self._ready.record(side)
        if not torch.cuda.is_current_stream_capturing():
            buf.record_stream(side)
        self.buf = buf

    def start_prefill(self):
        """Fork a prefill that was requested while `defer` was set.  The fill keeps one small CTA on every SM for its whole
        0.28 ms; a kernel that wants (nearly) all of an SM's shared memory cannot start beside it, so the caller forks the
        fill only after the last such kernel that follows the projection closely (the query selection's top-k, which keeps
        a row of scores in shared memory).  A request that is never started costs nothing: the backward then zero-fills
        the buffer itself."""
        req, self._pending, self.defer = self._pending, None, False
        if req is not None:
            self.prefill(*req)

    def join(self):
        """Make the current stream wait for a pending prefill (no-op otherwise)."""
        ev, self._ready = self._ready, None
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def take(self):
        """-> (buffer or None, written column offsets); resets the arena for the next pass."""
        self.join()
        buf, self.buf, self.base = self.buf, None, None
        written, self.written = self.written, set()
        return buf, written

    def grad_buffer(self, like):
        self.join()
        if self.buf is None:
            if self.grad_dtype is not None and self.grad_dtype != like.dtype:
                like = torch.empty(like.shape, dtype=self.grad_dtype, device=like.device)
            self.buf = _lib.zeros_like_fast(like)
        return self.buf

    def claim(self, col):
        """Each layer's value view may feed ONE sampler per backward pass: its gradient is accumulated in place, so a
        second consumer (or a second backward over a retained graph) would be added to a buffer autograd also sums."""
        if self.base is None:
            raise RuntimeError("tamtr_b200: the batched value projection was already back-propagated "
                               "(retain_graph / a second backward is not supported on the shared gradient arena)")
        if col in self.written:
            raise RuntimeError("tamtr_b200: a projected value view was consumed by two samplers; each view returned by "
                               "project_values() may be used once")
        self.written.add(col)


class _ValueProjFn(torch.autograd.Function):
    """value_all = feats @ W_cat^T + b_cat, returned as n per-layer head-major views [B, Lv, H, Dh]
    (transformer.py:273 for all layers at once).  Library GEMMs; the backward consumes the arena."""

    @staticmethod
    def forward(ctx, feats, w_cat, b_cat, arena, n, H):
        lp = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else feats.dtype
        B, Lv, C = feats.shape
        f2 = feats.reshape(B * Lv, C).to(lp)
        w = w_cat.to(lp)
        value_all = torch.addmm(b_cat.to(lp), f2, w.t()).view(B, Lv, -1)
        d = value_all.shape[-1] // n
        ctx.save_for_backward(f2, w)
        ctx.arena, ctx.n, ctx.d = arena, n, d
        ctx.meta = (value_all.shape, value_all.dtype, value_all.device, feats.dtype, w_cat.dtype, b_cat.dtype, feats.shape)
        ctx.set_materialize_grads(False)
        arena.base = value_all
        if any(ctx.needs_input_grad[:3]):
            arena.prefill(value_all.shape, value_all.dtype, value_all.device)
        return tuple(value_all[:, :, i * d:(i + 1) * d].view(B, Lv, H, d // H) for i in range(n))

    @staticmethod
    def backward(ctx, *grads):
        f2, w = ctx.saved_tensors
        arena, n, d = ctx.arena, ctx.n, ctx.d
        shape, dtype, device, fdt, wdt, bdt, fshape = ctx.meta
        buf, written = arena.take()
        if buf is None:
            buf = _lib.zeros_like_fast(torch.empty(shape, dtype=dtype, device=device))
        from_arena = True
        for i, g in enumerate(grads):
            if g is None:
                continue
            sl = buf[:, :, i * d:(i + 1) * d]
            if i * d in written:
                # accumulated in place by this layer's sampler backward, which hands autograd a stride-0 zero as a
                # token; anything else means autograd summed it with another consumer's gradient
                if any(st != 0 for st in g.stride()):
                    raise RuntimeError("tamtr_b200: a projected value view has a consumer besides its sampler")
                continue
            sl.add_(g.reshape(shape[0], shape[1], d))   # gradient did not come through the arena: add it
            from_arena = False
        g2 = buf.view(-1, shape[-1])
        if g2.dtype != dtype:                           # fp32 arena: one cast pass in front of the two GEMMs
            g2 = g2.to(dtype)
        grad_feats = (g2 @ w).view(fshape).to(fdt) if ctx.needs_input_grad[0] else None
        grad_w = None
        if ctx.needs_input_grad[1]:
            grad_w = (g2.t() @ f2 if dtype == torch.float32 else torch.mm(g2.t(), f2, out_dtype=torch.float32)).to(wdt)
        grad_b = None
        if ctx.needs_input_grad[2]:
            if from_arena and all(i * d in arena.bias_grad or grads[i] is None for i in range(n)):
                parts = [arena.bias_grad.get(i * d) for i in range(n)]
                grad_b = torch.cat([p if p is not None else torch.zeros(d, device=device) for p in parts]).to(bdt)
            else:
                grad_b = buf.view(-1, shape[-1]).float().sum(0).to(bdt)
        arena.bias_grad = {}
        return grad_feats, grad_w, grad_b, None, None, None


def project_values(feats, w_cat, b_cat, arena, n_layers, n_heads):
    """One GEMM for the value projections of all decoder layers -> n_layers views [B, Lv, H, Dh] that share `arena`."""
    return _ValueProjFn.apply(feats, w_cat, b_cat, arena, n_layers, n_heads)


class _MSDeformAttnFn(torch.autograd.Function):
    """value [B,Lv,H,Dh] (f32|bf16), loc [B,Lq,H,L,P,2] f32, attn [B,Lq,H,L,P] f32 -> out [B,Lq,H*Dh].
    `points` (tuple of per-level point counts, or None): the ragged variants of utils.py:92-191, where loc is
    [B,Lq,H,sum(points),2] and attn any [B,Lq,H,...] with sum(points) entries per head."""

    @staticmethod
    def forward(ctx, value, loc, attn, shapes, arena, points=None):
        _lib.require_cuda(value, loc, attn)
        value, tok_stride = _value_layout(value)
        loc = loc.contiguous().float()
        attn = attn.contiguous().float()
        B, Lv, H, Dh = value.shape
        sh, nl = _lib.shapes_array(shapes)
        Lq = loc.shape[1]
        out = torch.empty(B, Lq, H * Dh, dtype=value.dtype, device=value.device)
        if points is None:
            _, _, _, L, P, _ = loc.shape
            if nl != L:
                raise RuntimeError(f"tamtr_b200: {nl} level shapes given but sampling_locations has {L} levels")
            with _with_device(value):
                rc = _lib.lib().tamtr_msda_forward(value.data_ptr(), loc.data_ptr(), attn.data_ptr(), out.data_ptr(),
                                                   _lib.dtype_code(value), B, Lv, H, Dh, Lq, L, P, sh, tok_stride,
                                                   _lib.stream_ptr(value.device))
        else:
            pts = _points_array(points, nl, loc, attn)
            with _with_device(value):
                rc = _lib.lib().tamtr_msda_forward_ragged(value.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                                          out.data_ptr(), _lib.dtype_code(value), B, Lv, H, Dh, Lq, nl,
                                                          pts, sh, tok_stride, _lib.stream_ptr(value.device))
        _lib.check(rc, "msda_forward")
        ctx.points = None if points is None else tuple(int(p) for p in points)
        ctx.save_for_backward(value, loc, attn)
        ctx.shapes = [list(map(int, s)) for s in (shapes.tolist() if isinstance(shapes, torch.Tensor) else shapes)]
        ctx.tok_stride = tok_stride
        ctx.arena = arena if (arena is not None and tok_stride != H * Dh) else None
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        value, loc, attn = ctx.saved_tensors
        grad_out = grad_out.contiguous().to(value.dtype)
        B, Lv, H, Dh = value.shape
        Lq = loc.shape[1]
        sh, nl = _lib.shapes_array(ctx.shapes)
        tok_stride = ctx.tok_stride
        wsum = None
        if ctx.arena is not None:
            # accumulate into this layer's column slice of the shared, already zeroed buffer
            base = ctx.arena.base
            off = value.storage_offset() - (0 if base is None else base.storage_offset())
            ctx.arena.claim(off % value.stride(1))
            buf = ctx.arena.grad_buffer(base)
            grad_value = buf.view(-1)[off:].as_strided(value.shape, value.stride())
            zero = 0
            wsum = torch.empty(B, Lq, H, dtype=torch.float32, device=value.device)
        elif tok_stride != H * Dh:
            grad_value = torch.empty(B, Lv, H, Dh, dtype=value.dtype, device=value.device)
            tok_stride, zero = H * Dh, 1
            value = value.contiguous()
        else:
            grad_value = torch.empty_like(value)          # zeroed by the C call
            zero = 1
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
            if ctx.points is None:
                L, P = loc.shape[3], loc.shape[4]
                rc = _lib.lib().tamtr_msda_backward(grad_out.data_ptr(), value.data_ptr(), loc.data_ptr(),
                                                    attn.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
                                                    grad_attn.data_ptr(), _lib.dtype_code(value), B, Lv, H, Dh, Lq, L, P,
                                                    sh, tok_stride, zero, wsum.data_ptr() if wsum is not None else None,
                                                    _lib.dtype_code(grad_value), _lib.stream_ptr(value.device))
            else:
                pts = _points_array(ctx.points, nl, loc, attn)
                rc = _lib.lib().tamtr_msda_backward_ragged(grad_out.data_ptr(), value.data_ptr(), loc.data_ptr(),
                                                           attn.data_ptr(), grad_value.data_ptr(), grad_loc.data_ptr(),
                                                           grad_attn.data_ptr(), _lib.dtype_code(value), B, Lv, H, Dh, Lq,
                                                           nl, pts, sh, tok_stride, zero,
                                                           wsum.data_ptr() if wsum is not None else None,
                                                           _lib.dtype_code(grad_value), _lib.stream_ptr(value.device))
        _lib.check(rc, "msda_backward")
        if wsum is not None:      # value_proj bias gradient of this layer: sum_q wsum[q,h] * grad_out[q,h,:]
            ctx.arena.bias_grad[off % value.stride(1)] = torch.einsum(
                "bqh,bqhc->hc", wsum, grad_out.view(B, Lq, H, Dh).float()).reshape(-1)
        if ctx.arena is not None:
            # the gradient already sits in the arena (possibly in another dtype): autograd only needs a token
            grad_value = torch.zeros((), dtype=value.dtype, device=value.device).expand(value.shape)
        return grad_value, grad_loc, grad_attn, None, None, None


def _points_array(points, n_levels, loc, attn):
    """int32[L] host array of per-level point counts, checked against the tensors' sample counts."""
    import ctypes
    pts = [int(p) for p in points]
    if len(pts) != n_levels:
        raise RuntimeError(f"tamtr_b200: {len(pts)} point counts given for {n_levels} levels")
    S = sum(pts)
    if loc.dim() != 5 or loc.shape[3] != S or loc.shape[4] != 2:
        raise RuntimeError(f"tamtr_b200: sampling_locations {tuple(loc.shape)} must be [B, Lq, H, {S}, 2]")
    if attn.shape[:3] != loc.shape[:3] or attn[0, 0, 0].numel() != S:
        raise RuntimeError(f"tamtr_b200: attention_weights {tuple(attn.shape)} must hold {S} weights per (query, head)")
    return (ctypes.c_int32 * n_levels)(*pts)


def ms_deform_attn_ragged(value, value_spatial_shapes, sampling_locations, attention_weights, points):
    """The core op with a different number of points on each level: `points[l]` consecutive samples of
    sampling_locations [B,Lq,H,sum(points),2] belong to level l (the torch.split of utils.py:108 / :159)."""
    _lib.require_cuda(value, sampling_locations, attention_weights)
    if value.dtype == torch.float16 or value.dtype == torch.float64:
        out = _MSDeformAttnFn.apply(value.float(), sampling_locations, attention_weights, value_spatial_shapes, None,
                                    tuple(points))
        return out.to(value.dtype)
    return _MSDeformAttnFn.apply(value, sampling_locations, attention_weights, value_spatial_shapes, None, tuple(points))


def ms_deform_attn_cls(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Drop-in for multi_scale_deformable_attn_pytorch_cls (utils.py:92-140): 2 / 4 / 6 points on the three levels."""
    return ms_deform_attn_ragged(value, value_spatial_shapes, sampling_locations, attention_weights, (2, 4, 6))


def ms_deform_attn_box(value, value_spatial_shapes, sampling_locations, attention_weights):
    """Drop-in for multi_scale_deformable_attn_pytorch_box (utils.py:143-191): 6 / 4 / 2 points on the three levels."""
    return ms_deform_attn_ragged(value, value_spatial_shapes, sampling_locations, attention_weights, (6, 4, 2))


def ms_deform_attn(value, value_spatial_shapes, sampling_locations, attention_weights, arena=None):
    """Drop-in for multi_scale_deformable_attn_pytorch (utils.py:42).  fp16 values are computed in fp32.
    `arena`: optional ops.ValueArena when `value` is one of the views returned by split_values()."""
    _lib.require_cuda(value, sampling_locations, attention_weights)
    if value.dtype == torch.float16 or value.dtype == torch.float64:
        out = _MSDeformAttnFn.apply(value.float(), sampling_locations, attention_weights, value_spatial_shapes, None)
        return out.to(value.dtype)
    return _MSDeformAttnFn.apply(value, sampling_locations, attention_weights, value_spatial_shapes, arena)


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


def ms_deform_attn_corners_ragged(sampling_locations, value_spatial_shapes, points):
    """Index-math export for the ragged variants: (x0, y0) int32 [B,Lq,H,S], in-bounds flags uint8 [B,Lq,H,S,4]."""
    import ctypes
    _lib.require_cuda(sampling_locations)
    loc = sampling_locations.contiguous().float()
    B, Lq, H, S, _ = loc.shape
    sh, nl = _lib.shapes_array(value_spatial_shapes)
    pts = [int(p) for p in points]
    if len(pts) != nl or sum(pts) != S:
        raise RuntimeError(f"tamtr_b200: points {pts} do not describe {S} samples on {nl} levels")
    x0 = torch.empty(B, Lq, H, S, dtype=torch.int32, device=loc.device)
    y0 = torch.empty_like(x0)
    inb = torch.empty(B, Lq, H, S, 4, dtype=torch.uint8, device=loc.device)
    with _with_device(loc):
        rc = _lib.lib().tamtr_msda_corners_ragged(loc.data_ptr(), x0.data_ptr(), y0.data_ptr(), inb.data_ptr(), B, Lq, H,
                                                  nl, (ctypes.c_int32 * nl)(*pts), sh, _lib.stream_ptr(loc.device))
    _lib.check(rc, "msda_corners_ragged")
    return x0, y0, inb


# ------------------------------------------------------------------------------------ offsets / weights projection
def _proj_gemm(q2, w_cat):
    """[M,C] x [N,C]^T -> fp32 [M,N].  fp32 inputs: fp32 GEMM; 16-bit inputs: tensor-core GEMM with fp32 accumulate
    and fp32 OUTPUT, so locations never pass through a 16-bit rounding."""
    if q2.dtype == torch.float32:
        return q2 @ w_cat.t()
    return torch.mm(q2, w_cat.t(), out_dtype=torch.float32)


FUSED_PROJECTION = True     # bf16 activations: tamtr_locw_tc_forward (set False to force the GEMM + epilogue pair)


class _LocWFn(torch.autograd.Function):
    """q2 [M,C], sampling_offsets (w_off [2*H*S, C], b_off), attention_weights (w_attn [H*S, C], b_attn), ref [M,RL,RD]
    -> loc [M,H,L,P,2], attn [M,H,L,P] (fp32).

    forward  = bf16: ONE tcgen05 kernel, projections + epilogue (tamtr_locw_tc_forward, Kernel 3 fused);
               otherwise one library GEMM (both Linears of transformer.py:278-279 share their input) + tamtr_locw_forward
    backward = tamtr_locw_backward + library GEMMs."""

    @staticmethod
    def forward(ctx, q2, w_off, b_off, w_attn, b_attn, ref, shapes, H, L, P):
        lp = q2.dtype if q2.dtype in (torch.bfloat16, torch.float16) else torch.float32
        if torch.is_autocast_enabled("cuda"):
            lp = torch.get_autocast_dtype("cuda")
        q2c = q2.contiguous().to(lp)
        ref32 = ref.contiguous().float()
        M, C = q2c.shape
        RL, RD = ref32.shape[-2], ref32.shape[-1]
        if RD not in (2, 4):
            raise ValueError(f"Last dim of reference_points must be 2 or 4, but got {RD}.")  # transformer.py:295
        dev = q2c.device
        n_off = w_off.shape[0]
        need_ref = ctx.needs_input_grad[5]
        loc = torch.empty(M, H, L, P, 2, dtype=torch.float32, device=dev)
        attn = torch.empty(M, H, L, P, dtype=torch.float32, device=dev)
        lib = _lib.lib()
        bias = None
        if FUSED_PROJECTION and lp == torch.bfloat16 and lib.tamtr_locw_tc_supported(M, C, H, L, P, RL, RD):
            # Kernel 3 fused: the two weight matrices and biases are read where they are (no concatenation); `raw` (and
            # the concatenated bias) only exist if the reference boxes will need a gradient
            wo, wa = w_off.to(lp).contiguous(), w_attn.to(lp).contiguous()
            if b_off.dtype != b_attn.dtype or b_off.dtype not in (torch.float32, torch.bfloat16):
                b_off, b_attn = b_off.float(), b_attn.float()
            bo, ba = b_off.contiguous(), b_attn.contiguous()
            raw = torch.empty(M, 3 * H * L * P, dtype=torch.float32, device=dev) if need_ref else None
            with _with_device(q2c):
                rc = lib.tamtr_locw_tc_forward(q2c.data_ptr(), wo.data_ptr(), wa.data_ptr(), bo.data_ptr(), ba.data_ptr(),
                                               _lib.dtype_code(bo), ref32.data_ptr(), loc.data_ptr(), attn.data_ptr(),
                                               None if raw is None else raw.data_ptr(), M, C, H, L, P, RL, RD,
                                               _lib.stream_ptr(dev))
            _lib.check(rc, "locw_tc_forward")
            if need_ref:
                bias = torch.cat([bo, ba], 0).float()
        else:
            w_c = torch.cat([w_off, w_attn], 0).to(lp)
            bias = torch.cat([b_off, b_attn], 0).float()
            with torch.autocast("cuda", enabled=False):     # the epilogue kernel needs the fp32 GEMM output as is
                raw = _proj_gemm(q2c, w_c)
            assert raw.dtype == torch.float32
            sh, _ = _lib.shapes_array(shapes)
            with _with_device(raw):
                rc = lib.tamtr_locw_forward(raw.data_ptr(), bias.data_ptr(), ref32.data_ptr(), loc.data_ptr(),
                                            attn.data_ptr(), M, H, L, P, RL, RD, sh, _lib.stream_ptr(dev))
            _lib.check(rc, "locw_forward")
            wo, wa = w_c[:n_off], w_c[n_off:]
        ctx.save_for_backward(q2c, wo, wa, raw, bias, ref32, attn)
        ctx.dims = (M, H, L, P, RL, RD)
        ctx.shapes = [list(map(int, s)) for s in shapes]
        ctx.in_dtypes = (q2.dtype, w_off.dtype, b_off.dtype, w_attn.dtype, b_attn.dtype, ref.dtype)
        return loc, attn

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_loc, grad_attn):
        q2c, wo, wa, raw, bias, ref32, attn = ctx.saved_tensors
        M, H, L, P, RL, RD = ctx.dims
        grad_loc = grad_loc.contiguous().float()
        grad_attn = grad_attn.contiguous().float()
        grad_raw = torch.empty(M, 3 * H * L * P, dtype=torch.float32, device=attn.device)
        need_ref = ctx.needs_input_grad[5]
        grad_ref = torch.empty_like(ref32) if need_ref else None
        sh, _ = _lib.shapes_array(ctx.shapes)
        with _with_device(attn):
            rc = _lib.lib().tamtr_locw_backward(grad_loc.data_ptr(), grad_attn.data_ptr(), attn.data_ptr(),
                                                None if raw is None else raw.data_ptr(),
                                                None if bias is None else bias.data_ptr(),
                                                ref32.data_ptr(), grad_raw.data_ptr(),
                                                grad_ref.data_ptr() if need_ref else None,
                                                M, H, L, P, RL, RD, sh, _lib.stream_ptr(attn.device))
        _lib.check(rc, "locw_backward")
        qd, wod, bod, wad, bad, rd = ctx.in_dtypes
        n_off = wo.shape[0]
        g_lp = grad_raw.to(wo.dtype)
        grad_q = None
        if ctx.needs_input_grad[0]:
            grad_q = (g_lp[:, :n_off] @ wo).addmm_(g_lp[:, n_off:], wa).to(qd)
        grad_wo = grad_wa = grad_bo = grad_ba = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
            grad_w = (g_lp.t() @ q2c if wo.dtype == torch.float32 else torch.mm(g_lp.t(), q2c, out_dtype=torch.float32))
            grad_wo, grad_wa = grad_w[:n_off].to(wod), grad_w[n_off:].to(wad)
        if ctx.needs_input_grad[2] or ctx.needs_input_grad[4]:
            grad_b = col_sum(grad_raw)
            grad_bo, grad_ba = grad_b[:n_off].to(bod), grad_b[n_off:].to(bad)
        return (grad_q, grad_wo, grad_bo, grad_wa, grad_ba, (grad_ref.to(rd) if need_ref else None), None, None, None,
                None)


def sampling_locations_and_weights(query, refer_bbox, w_off, b_off, w_attn, b_attn, value_shapes, n_heads, n_levels,
                                   n_points):
    """transformer.py:278-293: bf16 activations -> one tcgen05 kernel (projections + epilogue); otherwise one GEMM + one
    epilogue kernel.

    query [B,Lq,C]; refer_bbox [B,Lq,RL,2|4]; returns loc [B,Lq,H,L,P,2] and attn [B,Lq,H,L,P], both fp32
    (index math and softmax stay fp32 whatever the activation dtype)."""
    _lib.require_cuda(query, refer_bbox)
    B, Lq, C = query.shape
    ref = refer_bbox.reshape(B * Lq, refer_bbox.shape[-2], refer_bbox.shape[-1])
    loc, attn = _LocWFn.apply(query.reshape(B * Lq, C), w_off, b_off, w_attn, b_attn, ref, value_shapes, n_heads, n_levels,
                              n_points)
    return (loc.view(B, Lq, n_heads, n_levels, n_points, 2), attn.view(B, Lq, n_heads, n_levels, n_points))


# ------------------------------------------------------------------------------------ box refinement
class _BoxRefineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, bbox, ref, eps):
        b32 = bbox.contiguous().float()
        r32 = ref.contiguous().float()
        out = torch.empty_like(b32)
        with _with_device(b32):
            rc = _lib.lib().tamtr_box_refine_forward(b32.data_ptr(), r32.data_ptr(), out.data_ptr(), b32.numel(),
                                                     float(eps), _lib.stream_ptr(b32.device))
        _lib.check(rc, "box_refine_forward")
        ctx.save_for_backward(out, r32)
        ctx.eps = float(eps)
        ctx.dtypes = (bbox.dtype, ref.dtype)
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        out, r32 = ctx.saved_tensors
        g = g.contiguous().float()
        gb = torch.empty_like(out)
        gr = torch.empty_like(out) if ctx.needs_input_grad[1] else None
        with _with_device(out):
            rc = _lib.lib().tamtr_box_refine_backward(g.data_ptr(), out.data_ptr(), r32.data_ptr(), gb.data_ptr(),
                                                      gr.data_ptr() if gr is not None else None, out.numel(), ctx.eps,
                                                      _lib.stream_ptr(out.device))
        _lib.check(rc, "box_refine_backward")
        return gb.to(ctx.dtypes[0]), (gr.to(ctx.dtypes[1]) if gr is not None else None), None


def box_refine(bbox, ref, eps=1e-5):
    """sigmoid(bbox + inverse_sigmoid(ref)) (transformer.py:875; utils.py:34-39) as one kernel; fp32 result."""
    _lib.require_cuda(bbox, ref)
    return _BoxRefineFn.apply(bbox, ref, eps)


# ------------------------------------------------------------------------------------ contrastive head
class _ContrastiveFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, logit_scale, bias):
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        w32 = w.contiguous().float()
        ls = logit_scale.detach().reshape(1).float().contiguous()
        bs = bias.detach().reshape(1).float().contiguous()
        B, Lq, C = x.shape
        K = w32.shape[1]
        out = torch.empty(B, Lq, K, dtype=torch.float32, device=x.device)
        with _with_device(x):
            rc = _lib.lib().tamtr_contrastive_forward(x.data_ptr(), w32.data_ptr(), ls.data_ptr(), bs.data_ptr(),
                                                      out.data_ptr(), _lib.dtype_code(x), B, Lq, K, C,
                                                      _lib.stream_ptr(x.device))
        _lib.check(rc, "contrastive_forward")
        ctx.save_for_backward(x, w32, ls, bs, out)
        ctx.w_dtype = w.dtype
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x, w32, ls, bs, out = ctx.saved_tensors
        B, Lq, C = x.shape
        K = w32.shape[1]
        grad_out = grad_out.contiguous().float()
        grad_x = torch.empty_like(x)
        scal = torch.empty(2, dtype=torch.float32, device=x.device)
        with _with_device(x):
            rc = _lib.lib().tamtr_contrastive_backward(grad_out.data_ptr(), x.data_ptr(), w32.data_ptr(),
                                                       ls.data_ptr(), grad_x.data_ptr(), scal.data_ptr(),
                                                       _lib.dtype_code(x), B, Lq, K, C, _lib.stream_ptr(x.device))
        _lib.check(rc, "contrastive_backward")
        grad_w = None
        if ctx.needs_input_grad[1]:
            # Text embeddings come from a frozen text model in TAM-TR (rtdetrworld/train.py:147-152), so this branch
            # is off the hot path; it is a [K,Lq]x[Lq,C] library GEMM per image plus the normalisation Jacobian.
            xh = torch.nn.functional.normalize(x.float(), dim=-1, eps=1e-12)
            wn = w32.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            wh = w32 / wn
            g_wh = torch.einsum("bqk,bqc->bkc", grad_out * ls.exp(), xh)
            grad_w = ((g_wh - wh * (wh * g_wh).sum(-1, keepdim=True)) / wn).to(ctx.w_dtype)
        return grad_x, grad_w, scal[0].reshape(()), scal[1].reshape(1)


def contrastive_head(x, w, logit_scale, bias):
    """block.py:534-541: x [B,Lq,C], w [B,K,C] -> [B,Lq,K]."""
    _lib.require_cuda(x, w)
    out = _ContrastiveFn.apply(x, w, logit_scale, bias)
    return out if x.dtype == torch.float32 else out.to(x.dtype)


# ------------------------------------------------------------------------------------ max-sigmoid gate
def _gate_uses_tensor_cores(embed, hc, HW, N, use_tensor_cores):
    ok = embed.dtype == torch.bfloat16 and hc == 32 and HW % 8 == 0 and N <= 128 and embed.data_ptr() % 16 == 0
    if use_tensor_cores and not ok:
        raise RuntimeError("tamtr_b200: tensor-core max-sigmoid gate needs bf16 activations, hc == 32, H*W % 8 == 0, "
                           "N <= 128")
    return ok if use_tensor_cores is None else bool(use_tensor_cores)


class _MaxSigmoidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, embed, guide, bias, nh, use_tensor_cores=None):
        embed = embed.contiguous()
        if embed.dtype not in (torch.float32, torch.bfloat16):
            embed = embed.float()
        B, C, Hh, Ww = embed.shape
        hc = C // nh
        g32 = guide.contiguous().float()                 # [B, N, nh, hc]
        N = g32.shape[1]
        b32 = bias.contiguous().float()
        aw = torch.empty(B, nh, Hh, Ww, dtype=torch.float32, device=embed.device)
        amax = torch.empty(B, nh, Hh, Ww, dtype=torch.uint8, device=embed.device)
        with _with_device(embed):
            if _gate_uses_tensor_cores(embed, hc, Hh * Ww, N, use_tensor_cores):
                rc = _lib.lib().tamtr_max_sigmoid_tc_forward(embed.data_ptr(), g32.data_ptr(), b32.data_ptr(),
                                                             aw.data_ptr(), amax.data_ptr(), B, nh, hc, Hh * Ww, N,
                                                             _lib.stream_ptr(embed.device))
            else:
                rc = _lib.lib().tamtr_max_sigmoid_forward(embed.data_ptr(), g32.data_ptr(), b32.data_ptr(),
                                                          aw.data_ptr(), amax.data_ptr(), _lib.dtype_code(embed), B,
                                                          nh, hc, Hh * Ww, N, _lib.stream_ptr(embed.device))
        _lib.check(rc, "max_sigmoid_forward")
        ctx.save_for_backward(embed, g32, aw, amax)
        ctx.meta = (B, nh, hc, Hh * Ww, N, guide.dtype, bias.dtype)
        return aw

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_aw):
        embed, g32, aw, amax = ctx.saved_tensors
        B, nh, hc, HW, N, gdt, bdt = ctx.meta
        grad_aw = grad_aw.contiguous().float()
        grad_embed = torch.empty_like(embed)
        grad_guide = torch.empty_like(g32)
        grad_bias = torch.empty(nh, dtype=torch.float32, device=embed.device)
        with _with_device(embed):
            rc = _lib.lib().tamtr_max_sigmoid_backward(grad_aw.data_ptr(), aw.data_ptr(), amax.data_ptr(),
                                                       embed.data_ptr(), g32.data_ptr(), grad_embed.data_ptr(),
                                                       grad_guide.data_ptr(), grad_bias.data_ptr(),
                                                       _lib.dtype_code(embed), B, nh, hc, HW, N,
                                                       _lib.stream_ptr(embed.device))
        _lib.check(rc, "max_sigmoid_backward")
        return grad_embed, grad_guide.to(gdt), grad_bias.to(bdt), None, None


def max_sigmoid_gate(embed, guide, bias, nh, use_tensor_cores=None):
    """extra_modules/block.py:216-220: embed [B,nh*hc,H,W], guide [B,N,nh,hc], bias [nh] -> aw [B,nh,H,W] fp32.
    bf16 activations with hc == 32 run on the tcgen05 kernel (csrc/maxsig_tc.cu); fp32 ones on the exact CUDA-core
    kernel.  `use_tensor_cores` forces the choice (tests / benchmarks)."""
    _lib.require_cuda(embed, guide, bias)
    return _MaxSigmoidFn.apply(embed, guide, bias, nh, use_tensor_cores)


# ------------------------------------------------------------------------------------ Linear layers of the decoder
def _lowp(*tensors):
    """dtype a Linear computes in: the autocast dtype when autocast is on (as F.linear would), else the input's."""
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else tensors[0].dtype


def col_sum(g2):
    """sum over rows of g2 [rows, n] -> fp32 [n] (tamtr_col_sum); library reduction for shapes the kernel does not take."""
    rows, n = g2.shape
    np_ = 4 if g2.dtype == torch.float32 else 8
    if g2.dtype not in (torch.float32, torch.bfloat16) or n % np_ != 0 or g2.data_ptr() % 16 != 0 or not g2.is_contiguous():
        return g2.float().sum(0)
    out = torch.empty(n, dtype=torch.float32, device=g2.device)
    with _with_device(g2):
        _lib.check(_lib.lib().tamtr_col_sum(g2.data_ptr(), out.data_ptr(), _lib.dtype_code(g2), rows, n,
                                            _lib.stream_ptr(g2.device)), "col_sum")
    return out


class _LinearFn(torch.autograd.Function):
    """F.linear with the same arithmetic (library GEMMs, bias in the GEMM epilogue) whose backward computes the bias
    gradient with tamtr_col_sum instead of the library's generic reduction."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lp = _lowp(x)
        x2 = x.reshape(-1, x.shape[-1]).to(lp)
        w = weight.to(lp)
        y = torch.addmm(bias.to(lp), x2, w.t()) if bias is not None else x2 @ w.t()
        ctx.save_for_backward(x2, w)
        ctx.meta = (x.shape, x.dtype, weight.dtype, None if bias is None else bias.dtype)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x2, w = ctx.saved_tensors
        xshape, xdt, wdt, bdt = ctx.meta
        g2 = g.reshape(-1, g.shape[-1]).to(w.dtype).contiguous()
        dx = (g2 @ w).view(xshape).to(xdt) if ctx.needs_input_grad[0] else None
        dw = (g2.t() @ x2).to(wdt) if ctx.needs_input_grad[1] else None
        db = col_sum(g2).to(bdt) if (bdt is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db


def linear(x, layer):
    """`layer(x)` for an nn.Linear on CUDA tensors (transformer.py:166-176, 535-537, ...)."""
    if not x.is_cuda:
        return layer(x)
    return _LinearFn.apply(x, layer.weight, layer.bias)


class _LinearReluFn(torch.autograd.Function):
    """relu(F.linear(x)) with the ReLU in the GEMM epilogue (one launch instead of two; the pre-activation is never
    stored: the backward masks the incoming gradient with the output's sign)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lp = _lowp(x)
        x2 = x.reshape(-1, x.shape[-1]).to(lp)
        w = weight.to(lp)
        y = torch._addmm_activation(bias.to(lp), x2, w.t(), use_gelu=False)
        ctx.save_for_backward(x2, w, y)
        ctx.meta = (x.shape, x.dtype, weight.dtype, bias.dtype)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        x2, w, y = ctx.saved_tensors
        xshape, xdt, wdt, bdt = ctx.meta
        g2 = torch.ops.aten.threshold_backward(g.reshape(-1, g.shape[-1]).to(w.dtype), y, 0).contiguous()
        dx = (g2 @ w).view(xshape).to(xdt) if ctx.needs_input_grad[0] else None
        dw = (g2.t() @ x2).to(wdt) if ctx.needs_input_grad[1] else None
        db = col_sum(g2).to(bdt) if ctx.needs_input_grad[2] else None
        return dx, dw, db


def linear_relu(x, layer):
    """relu(layer(x)) for an nn.Linear (the hidden layers of transformer.py:162-176 MLP and of the decoder FFN, :609-619)."""
    if not x.is_cuda or layer.bias is None:
        return torch.relu(linear(x, layer))
    return _LinearReluFn.apply(x, layer.weight, layer.bias)


class _InProjFn(torch.autograd.Function):
    """Packed input projection of nn.MultiheadAttention when query and key share their input (decoder self-attention,
    transformer.py:544-547: q = k = embed + pos, v = embed): two GEMMs, gradients written into ONE [3d, d] / [3d]
    buffer so that in_proj_weight / in_proj_bias receive a single gradient tensor."""

    @staticmethod
    def forward(ctx, xqk, xv, weight, bias):
        lp = _lowp(xqk)
        d = weight.shape[1]
        a = xqk.reshape(-1, d).to(lp)
        c = xv.reshape(-1, d).to(lp)
        w = weight.to(lp)
        b = bias.to(lp)
        qk = torch.addmm(b[:2 * d], a, w[:2 * d].t())
        v = torch.addmm(b[2 * d:], c, w[2 * d:].t())
        ctx.save_for_backward(a, c, w)
        ctx.meta = (xqk.shape, xqk.dtype, xv.dtype, weight.dtype, bias.dtype)
        lead = xqk.shape[:-1]
        return qk.view(*lead, 2 * d), v.view(*lead, d)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gqk, gv):
        a, c, w = ctx.saved_tensors
        xshape, qdt, vdt, wdt, bdt = ctx.meta
        d = w.shape[1]
        gqk2 = gqk.reshape(-1, 2 * d).to(w.dtype).contiguous()
        gv2 = gv.reshape(-1, d).to(w.dtype).contiguous()
        dxqk = (gqk2 @ w[:2 * d]).view(xshape).to(qdt) if ctx.needs_input_grad[0] else None
        dxv = (gv2 @ w[2 * d:]).view(xshape).to(vdt) if ctx.needs_input_grad[1] else None
        dw = db = None
        if ctx.needs_input_grad[2]:
            dw = torch.empty_like(w)
            torch.mm(gqk2.t(), a, out=dw[:2 * d])
            torch.mm(gv2.t(), c, out=dw[2 * d:])
            dw = dw.to(wdt)
        if ctx.needs_input_grad[3]:
            db = torch.cat([col_sum(gqk2), col_sum(gv2)]).to(bdt)
        return dxqk, dxv, dw, db


class _SelfAttnFn(torch.autograd.Function):
    """softmax(q k^T / sqrt(Dh) + mask) v on the packed projections (csrc/selfattn.cu): qk [B, L, 2d] bf16, v [B, L, d] bf16,
    blocked = attention_mask_bits(mask) or None -> o [B, L, d] bf16.  Gradients come back in the same packed layouts."""

    @staticmethod
    def forward(ctx, qk, v, blocked, n_heads):
        B, L, d = v.shape
        qk, v = qk.contiguous(), v.contiguous()
        o = torch.empty_like(v)
        lse = torch.empty(B, n_heads, L, dtype=torch.float32, device=v.device)
        with _with_device(v):
            rc = _lib.lib().tamtr_self_attention_forward(qk.data_ptr(), v.data_ptr(), None if blocked is None else blocked.data_ptr(),
                                                         o.data_ptr(), lse.data_ptr(), B, L, n_heads, d // n_heads,
                                                         _lib.stream_ptr(v.device))
        _lib.check(rc, "self_attention_forward")
        ctx.save_for_backward(qk, v, o, lse, blocked)
        ctx.n_heads = n_heads
        return o

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, go):
        qk, v, o, lse, blocked = ctx.saved_tensors
        B, L, d = v.shape
        H = ctx.n_heads
        go = go.contiguous().to(v.dtype)
        d_qk, d_v = torch.empty_like(qk), torch.empty_like(v)
        lp = _lib.lib().tamtr_self_attention_padded_len(L)
        scratch = torch.empty(2, B, H, lp, lp, dtype=v.dtype, device=v.device)
        with _with_device(v):
            rc = _lib.lib().tamtr_self_attention_backward(qk.data_ptr(), v.data_ptr(), None if blocked is None else blocked.data_ptr(),
                                                          o.data_ptr(), go.data_ptr(), lse.data_ptr(), d_qk.data_ptr(),
                                                          d_v.data_ptr(), scratch.data_ptr(), B, L, H, d // H,
                                                          _lib.stream_ptr(v.device))
        _lib.check(rc, "self_attention_backward")
        return d_qk, d_v, None, None


def attention_mask_bits(attn_mask):
    """bool [L, L] (True = blocked) -> the bit-packed form the attention kernels read (int64 [L, ceil(L / 64)]), or None.
    Packed once per mask TENSOR: the decoder hands the same tensor to every layer of a forward."""
    if attn_mask is None:
        return None
    hit = getattr(attn_mask, "_tamtr_bits", None)
    if hit is not None and hit[0] == attn_mask._version:
        return hit[1]
    L = attn_mask.shape[0]
    u8 = attn_mask.to(torch.uint8).contiguous() if attn_mask.dtype != torch.uint8 else attn_mask.contiguous()
    bits = torch.empty(_lib.lib().tamtr_self_attention_mask_words(L), dtype=torch.int64, device=attn_mask.device)
    with _with_device(attn_mask):
        _lib.check(_lib.lib().tamtr_self_attention_pack_mask(u8.data_ptr(), bits.data_ptr(), L, _lib.stream_ptr(attn_mask.device)),
                   "self_attention_pack_mask")
    try:
        attn_mask._tamtr_bits = (attn_mask._version, bits)
    except Exception:
        pass
    return bits


FUSED_SELF_ATTENTION = True      # False: the library's scaled_dot_product_attention (kept for A/B tests)


def self_attention(mha, x_qk, x_v, attn_mask=None):
    """nn.MultiheadAttention(x_qk, x_qk, x_v, attn_mask=..., need_weights=False)[0] for batch-first [B, L, d] inputs
    (the reference feeds it sequence-first through two transposes, transformer.py:546): packed in-projection, attention,
    out-projection.  attn_mask: bool [L, L] with True = blocked (nn.MultiheadAttention's convention) or an additive float
    mask.  bf16 projections with head dimension 32 / 64 and a bool (or no) mask run on the fused kernels of csrc/selfattn.cu,
    which read q / k / v where the projection wrote them; anything else goes through F.scaled_dot_product_attention."""
    B, L, d = x_qk.shape
    H = mha.num_heads
    qk, v = _InProjFn.apply(x_qk, x_v, mha.in_proj_weight, mha.in_proj_bias)
    if (FUSED_SELF_ATTENTION and qk.dtype == torch.bfloat16 and v.dtype == torch.bfloat16 and d % H == 0
            and (attn_mask is None or (attn_mask.dtype == torch.bool and tuple(attn_mask.shape) == (L, L)))
            and (mha.dropout == 0.0 or not mha.training) and _lib.lib().tamtr_self_attention_supported(L, H, d // H)):
        return linear(_SelfAttnFn.apply(qk, v, attention_mask_bits(attn_mask), H), mha.out_proj)
    q = qk[..., :d].view(B, L, H, d // H).transpose(1, 2)
    k = qk[..., d:].view(B, L, H, d // H).transpose(1, 2)
    v = v.view(B, L, H, d // H).transpose(1, 2)
    if attn_mask is not None and attn_mask.dtype == torch.bool:
        attn_mask = attn_mask.logical_not()                            # SDPA: True = attend
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask,
                                                         dropout_p=mha.dropout if mha.training else 0.0)
    return linear(o.transpose(1, 2).reshape(B, L, d), mha.out_proj)


# ------------------------------------------------------------------------------------ residual add + LayerNorm
def add_layer_norm_supported(x, d):
    return x.is_cuda and d % 128 == 0 and 128 <= d <= 512 and x.dtype in (torch.float32, torch.bfloat16)


class _AddLayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, res, weight, bias, eps, track=True):
        d = x.shape[-1]
        xc = x.contiguous()
        rc_ = None if res is None else res.contiguous()
        if rc_ is not None and rc_.dtype not in (torch.float32, torch.bfloat16):
            rc_ = rc_.float()
        rows = xc.numel() // d
        out_dtype = torch.float32 if (xc.dtype == torch.float32 or (rc_ is not None and rc_.dtype == torch.float32)
                                      or torch.is_autocast_enabled("cuda")) else xc.dtype
        w32 = weight.detach().float().contiguous()
        b32 = bias.detach().float().contiguous()
        y = torch.empty(xc.shape, dtype=out_dtype, device=xc.device)
        # `track`: grad mode at the call site (needs_input_grad alone is True for parameters even under no_grad, which
        # made inference write z / mean / rstd for nothing)
        need = bool(track) and any(ctx.needs_input_grad[:4])
        # no residual and fp32 input: z IS x, nothing to write (the backward reads x)
        z_is_x = rc_ is None and xc.dtype == torch.float32
        z = torch.empty(xc.shape, dtype=torch.float32, device=xc.device) if (need and not z_is_x) else None
        mean = torch.empty(rows, dtype=torch.float32, device=xc.device) if need else None
        rstd = torch.empty(rows, dtype=torch.float32, device=xc.device) if need else None
        with _with_device(xc):
            rc = _lib.lib().tamtr_add_layernorm_forward(
                xc.data_ptr(), _lib.dtype_code(xc), None if rc_ is None else rc_.data_ptr(),
                0 if rc_ is None else _lib.dtype_code(rc_), w32.data_ptr(), b32.data_ptr(), y.data_ptr(),
                _lib.dtype_code(y), None if z is None else z.data_ptr(), None if mean is None else mean.data_ptr(),
                None if rstd is None else rstd.data_ptr(), rows, d, float(eps), _lib.stream_ptr(xc.device))
        _lib.check(rc, "add_layernorm_forward")
        if need:
            ctx.save_for_backward(xc if z_is_x else z, mean, rstd, w32)
            ctx.meta = (xc.dtype, None if rc_ is None else rc_.dtype, weight.dtype, bias.dtype, rows, d)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        z, mean, rstd, w32 = ctx.saved_tensors
        xdt, rdt, wdt, bdt, rows, d = ctx.meta
        gy = gy.contiguous()
        if gy.dtype not in (torch.float32, torch.bfloat16):
            gy = gy.float()
        dev = gy.device
        need_x, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1] and rdt is not None
        dx = torch.empty(z.shape, dtype=xdt, device=dev) if need_x else None
        # one tensor serves both inputs when they share a dtype
        if need_r and need_x and rdt == xdt:
            dres, dres_ptr = dx, None
        else:
            dres = torch.empty(z.shape, dtype=rdt, device=dev) if need_r else None
            dres_ptr = None if dres is None else dres.data_ptr()
        dwb = torch.empty(2, d, dtype=torch.float32, device=dev)       # zeroed by the call
        with _with_device(gy):
            rc = _lib.lib().tamtr_add_layernorm_backward(
                gy.data_ptr(), _lib.dtype_code(gy), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w32.data_ptr(),
                None if dx is None else dx.data_ptr(), 0 if dx is None else _lib.dtype_code(dx), dres_ptr,
                0 if dres_ptr is None else _lib.dtype_code(dres), dwb.data_ptr(), rows, d, _lib.stream_ptr(dev))
        _lib.check(rc, "add_layernorm_backward")
        return dx, dres, dwb[0].to(wdt), dwb[1].to(bdt), None, None


def add_layer_norm(x, res, norm):
    """`norm(x + res)` (res may be None) for an nn.LayerNorm over the last dimension: one kernel each way
    (csrc/layernorm.cu).  Output is fp32 when an input is fp32 or under autocast (autocast runs layer_norm in fp32),
    otherwise the input dtype."""
    _lib.require_cuda(x, res)
    return _AddLayerNormFn.apply(x, res, norm.weight, norm.bias, norm.eps, torch.is_grad_enabled())


class _AddLayerNormSidesFn(torch.autograd.Function):
    """norm(x + res) -> (y fp32, bf16(y) | None, bf16(y + pos) | None): the add + LayerNorm kernel also writes what the
    bf16 consumers of the normalised stream read, and its backward sums the gradients of all three outputs while it loads
    them (tamtr_add_layernorm_forward_sides / _backward_sides)."""

    @staticmethod
    def forward(ctx, x, res, weight, bias, eps, pos, want_lp, track):
        d = x.shape[-1]
        xc, rc_ = x.contiguous(), res.contiguous()
        pc = None if pos is None else pos.contiguous()
        rows = xc.numel() // d
        dev = xc.device
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        y = torch.empty(xc.shape, dtype=torch.float32, device=dev)
        y_lp = torch.empty(xc.shape, dtype=torch.bfloat16, device=dev) if want_lp else None
        q_lp = torch.empty(xc.shape, dtype=torch.bfloat16, device=dev) if pc is not None else None
        need = bool(track) and any(ctx.needs_input_grad[:4] + (ctx.needs_input_grad[5],))
        z = torch.empty(xc.shape, dtype=torch.float32, device=dev) if need else None
        mean = torch.empty(rows, dtype=torch.float32, device=dev) if need else None
        rstd = torch.empty(rows, dtype=torch.float32, device=dev) if need else None
        ptr = lambda t: None if t is None else t.data_ptr()                         # noqa: E731
        with _with_device(xc):
            rc = _lib.lib().tamtr_add_layernorm_forward_sides(
                xc.data_ptr(), _lib.dtype_code(xc), rc_.data_ptr(), _lib.dtype_code(rc_), w32.data_ptr(), b32.data_ptr(),
                y.data_ptr(), _lib.dtype_code(y), ptr(z), ptr(mean), ptr(rstd), ptr(pc),
                0 if pc is None else _lib.dtype_code(pc), ptr(y_lp), ptr(q_lp), rows, d, float(eps), _lib.stream_ptr(dev))
        _lib.check(rc, "add_layernorm_forward_sides")
        if need:
            ctx.save_for_backward(z, mean, rstd, w32)
            ctx.meta = (xc.dtype, rc_.dtype, weight.dtype, bias.dtype, None if pc is None else pc.dtype, rows, d)
        ctx.set_materialize_grads(False)
        return y, y_lp, q_lp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy, g_ylp, g_qlp):
        if gy is None and g_ylp is None and g_qlp is None:
            return (None,) * 8
        z, mean, rstd, w32 = ctx.saved_tensors
        xdt, rdt, wdt, bdt, pdt, rows, d = ctx.meta
        dev = z.device
        if gy is not None:
            gy = gy.contiguous()
            if gy.dtype not in (torch.float32, torch.bfloat16):
                gy = gy.float()
        g1 = None if g_ylp is None else g_ylp.contiguous().to(torch.bfloat16)
        g2 = None if g_qlp is None else g_qlp.contiguous().to(torch.bfloat16)
        need_x, need_r = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        dx = torch.empty(z.shape, dtype=xdt, device=dev) if need_x else None
        if need_r and need_x and rdt == xdt:
            dres, dres_ptr = dx, None
        else:
            dres = torch.empty(z.shape, dtype=rdt, device=dev) if need_r else None
            dres_ptr = None if dres is None else dres.data_ptr()
        dwb = torch.empty(2, d, dtype=torch.float32, device=dev)       # zeroed by the call
        ptr = lambda t: None if t is None else t.data_ptr()                         # noqa: E731
        with _with_device(z):
            rc = _lib.lib().tamtr_add_layernorm_backward_sides(
                ptr(gy), 0 if gy is None else _lib.dtype_code(gy), ptr(g1), ptr(g2), z.data_ptr(), mean.data_ptr(),
                rstd.data_ptr(), w32.data_ptr(), ptr(dx), 0 if dx is None else _lib.dtype_code(dx), dres_ptr,
                0 if dres_ptr is None else _lib.dtype_code(dres), dwb.data_ptr(), rows, d, _lib.stream_ptr(dev))
        _lib.check(rc, "add_layernorm_backward_sides")
        d_pos = None
        if ctx.needs_input_grad[5] and g2 is not None:
            d_pos = g2 if pdt == g2.dtype else g2.to(pdt)              # d(y + pos)/d(pos) = 1
        return dx, dres, dwb[0].to(wdt), dwb[1].to(bdt), None, d_pos, None, None


def add_layer_norm_sides(x, res, norm, pos=None, want_lp=False):
    """(norm(x + res) in fp32, its bf16 copy if `want_lp`, bf16(norm(x + res) + pos) if `pos` is given) -- one kernel each
    way.  For the decoder layers under bf16 autocast (modules._layer_forward_lowp)."""
    _lib.require_cuda(x, res)
    return _AddLayerNormSidesFn.apply(x, res, norm.weight, norm.bias, norm.eps, pos, bool(want_lp), torch.is_grad_enabled())


class _PosCastFn(torch.autograd.Function):
    """x -> (x, bf16(x), bf16(x + pos)): the operands of the self-attention projections (transformer.py:544-547) from one
    pass over x.  x itself is handed through so that the gradient of its other consumer (the residual connection) arrives
    at this node and is summed with the two bf16 gradients by one kernel."""

    @staticmethod
    def forward(ctx, x, pos):
        xc, pc = x.contiguous(), pos.contiguous()
        x_lp = torch.empty(xc.shape, dtype=torch.bfloat16, device=xc.device)
        q_lp = torch.empty(xc.shape, dtype=torch.bfloat16, device=xc.device)
        with _with_device(xc):
            rc = _lib.lib().tamtr_pos_cast(xc.data_ptr(), _lib.dtype_code(xc), pc.data_ptr(), _lib.dtype_code(pc),
                                           x_lp.data_ptr(), q_lp.data_ptr(), xc.numel(), _lib.stream_ptr(xc.device))
        _lib.check(rc, "pos_cast")
        ctx.meta = (x.dtype, pos.dtype, x.shape)
        ctx.set_materialize_grads(False)
        return x.view_as(x), x_lp, q_lp

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gx, g_xlp, g_qlp):
        xdt, pdt, shape = ctx.meta
        gs = [g.contiguous() for g in (gx, g_xlp, g_qlp) if g is not None]
        d_pos = None
        if ctx.needs_input_grad[1] and g_qlp is not None:
            d_pos = g_qlp if g_qlp.dtype == pdt else g_qlp.to(pdt)
        if not gs or not ctx.needs_input_grad[0]:
            return None, d_pos
        if len(gs) == 1 and gs[0].dtype == xdt:
            return gs[0], d_pos
        if any(g.dtype not in (torch.float32, torch.bfloat16) for g in gs):
            gs = [g.float() for g in gs]
        gs += [None] * (3 - len(gs))
        out = torch.empty(shape, dtype=xdt, device=gs[0].device)
        ptr = lambda t: None if t is None else t.data_ptr()                         # noqa: E731
        code = lambda t: 0 if t is None else _lib.dtype_code(t)                     # noqa: E731
        with _with_device(out):
            rc = _lib.lib().tamtr_grad_sum3(ptr(gs[0]), code(gs[0]), ptr(gs[1]), code(gs[1]), ptr(gs[2]), code(gs[2]),
                                            out.data_ptr(), _lib.dtype_code(out), out.numel(), _lib.stream_ptr(out.device))
        _lib.check(rc, "grad_sum3")
        return out, d_pos


def pos_cast(x, pos):
    """-> (x, bf16(x), bf16(x + pos)) with one kernel forward and one backward."""
    _lib.require_cuda(x, pos)
    return _PosCastFn.apply(x, pos)


def to_channels_last(x):
    """[B, C, H, W] -> the same logical tensor in channels-last memory ([B, H, W, C] storage).  No copy when the caller
    already holds channels-last maps; otherwise one pass of tamtr_nchw_to_nhwc."""
    _lib.require_cuda(x)
    B, C, H, W = x.shape
    if x.permute(0, 2, 3, 1).is_contiguous():
        return x
    x = x.contiguous()
    out = torch.empty((B, H, W, C), dtype=x.dtype, device=x.device)
    with _with_device(x):
        _lib.check(_lib.lib().tamtr_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), _lib.dtype_code(x), B, C, H * W,
                                                 _lib.stream_ptr(x.device)), "nchw_to_nhwc")
    return out.permute(0, 3, 1, 2)


def gate_conv3x3_supported(x, weight, nh):
    co, ci, kh, kw = weight.shape
    return (x.is_cuda and x.dtype == torch.bfloat16 and kh == 3 and kw == 3 and ci % 64 == 0 and co % 32 == 0
            and 32 <= co <= 256 and co % nh == 0 and (co // nh) % 32 == 0)


_OHWI_CACHE = {}


def _weight_ohwi(weight):
    """[Co, Ci, 3, 3] -> bf16 [Co, 3, 3, Ci] (the K-major B operand), cached per weight TENSOR OBJECT and version:
    inference calls the block with the same parameter every time.  The entry holds a weak reference and is only used
    when it still points at this very tensor (an address or id can be recycled by another tensor)."""
    import weakref
    hit = _OHWI_CACHE.get(id(weight))
    if hit is not None and hit[0]() is weight and hit[1] == weight._version:
        return hit[2]
    if len(_OHWI_CACHE) >= 64:
        for k in [k for k, v in _OHWI_CACHE.items() if v[0]() is None]:
            del _OHWI_CACHE[k]
        if len(_OHWI_CACHE) >= 64:
            _OHWI_CACHE.clear()
    out = weight.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    try:
        _OHWI_CACHE[id(weight)] = (weakref.ref(weight), weight._version, out)
    except TypeError:
        pass
    return out


def _gate_conv3x3_launch(x, weight, bn_scale, bn_shift, gate, nh):
    B, Ci, H, W = x.shape
    Co = weight.shape[0]
    x_cl = to_channels_last(x)                                        # storage [B, H, W, Ci]
    w_ohwi = _weight_ohwi(weight)
    s32 = bn_scale.detach().float().contiguous()
    t32 = bn_shift.detach().float().contiguous()
    g32 = None if gate is None else gate.detach().float().contiguous()
    y = torch.empty((B, H, W, Co), dtype=torch.bfloat16, device=x.device)
    with _with_device(x):
        rc = _lib.lib().tamtr_gate_conv3x3_tc_forward(x_cl.data_ptr(), w_ohwi.data_ptr(), s32.data_ptr(), t32.data_ptr(),
                                                      None if g32 is None else g32.data_ptr(), y.data_ptr(), B, H, W,
                                                      Ci, Co, nh, _lib.stream_ptr(x.device))
    _lib.check(rc, "gate_conv3x3_tc_forward")
    return x_cl, y.permute(0, 3, 1, 2)                                # logical [B, Co, H, W], channels-last memory


def gate_conv3x3(x, weight, bn_scale, bn_shift, gate, nh):
    """extra_modules/block.py:222-225 with BatchNorm folded to an affine: (conv3x3(x) * bn_scale + bn_shift) * gate,
    one tcgen05 kernel (csrc/gateconv_tc.cu).  x [B,Ci,H,W] bf16 (NCHW or channels-last), weight [Co,Ci,3,3],
    gate [B,nh,H,W] or None -> [B,Co,H,W] bf16 in channels-last memory.  Inference path: no autograd."""
    _lib.require_cuda(x, weight, bn_scale, bn_shift, gate)
    if not gate_conv3x3_supported(x, weight, nh):
        raise RuntimeError("tamtr_b200: gate_conv3x3 needs bf16 activations, a 3x3 kernel, Cin % 64 == 0, "
                           "Cout % 32 == 0, Cout <= 256 and (Cout / nh) % 32 == 0")
    return _gate_conv3x3_launch(x, weight, bn_scale, bn_shift, gate, nh)[1]


_IDENTITY_AFFINE = {}


def _identity_affine(co, device):
    key = (co, device)
    if key not in _IDENTITY_AFFINE:
        _IDENTITY_AFFINE[key] = (torch.ones(co, dtype=torch.float32, device=device),
                                 torch.zeros(co, dtype=torch.float32, device=device))
    return _IDENTITY_AFFINE[key]


class _Conv3x3TcFn(torch.autograd.Function):
    """The raw 3x3 convolution (stride 1, pad 1, no bias) on the tcgen05 kernel, forward AND data gradient: dgrad of a
    stride-1 / pad-1 3x3 convolution is the same convolution of grad_out with the filter rotated by 180 degrees and its
    channel roles swapped (w'[ci, co, ky, kx] = w[co, ci, 2 - ky, 2 - kx]), so it runs on the same implicit-GEMM kernel
    whenever the swapped shape is one the kernel takes.  The weight gradient (a GEMM whose K is the pixel dimension) is a
    library call."""

    @staticmethod
    def forward(ctx, x, weight):
        one, zero = _identity_affine(weight.shape[0], x.device)
        x_cl, y = _gate_conv3x3_launch(x, weight, one, zero, None, 1)
        ctx.save_for_backward(x_cl, weight)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x_cl, weight = ctx.saved_tensors
        gy = gy.to(x_cl.dtype)
        gx = gw = None
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if need_x:
            w_d = weight.detach().flip(2, 3).transpose(0, 1)          # [Ci, Co, 3, 3]: rotated, channel roles swapped
            if gate_conv3x3_supported(gy, w_d, 1):
                one, zero = _identity_affine(w_d.shape[0], gy.device)
                gx = _gate_conv3x3_launch(gy, w_d, one, zero, None, 1)[1]
                need_x = False
        if need_x or need_w:
            gx_l, gw, _ = torch.ops.aten.convolution_backward(
                gy, x_cl, weight.to(x_cl.dtype), None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [need_x, need_w, False])
            gx = gx_l if need_x else gx
        return gx, None if gw is None else gw.to(weight.dtype)


def conv3x3_tc(x, weight):
    """Conv2d(k=3, s=1, p=1, bias=False) of `proj_conv` for bf16 activations (training path: BatchNorm statistics and
    the gate multiply stay outside)."""
    _lib.require_cuda(x, weight)
    if not gate_conv3x3_supported(x, weight, 1):
        raise RuntimeError("tamtr_b200: conv3x3_tc needs bf16 activations, Cin % 64 == 0, Cout % 32 == 0, Cout <= 256")
    return _Conv3x3TcFn.apply(x, weight)


# ------------------------------------------------------------------------------------ sparse-gradient plumbing
class GradHub:
    """Collects row-sparse gradients for one activation so that they are added IN PLACE to the dense gradient that
    the other consumers produce, instead of each materialising a mostly-zero tensor of the activation's size.

    Used for `feats` [B, Lv, d] in the detection heads: the decoder layers send it a dense gradient (value_proj),
    the query-selection branch only touches B*nq of its B*Lv rows (head.py:1233-1254 gathers top-k rows)."""

    def __init__(self):
        self.pending = []


class _HubFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, hub):
        ctx.hub = hub
        ctx.set_materialize_grads(False)
        ctx.meta = (x.shape, x.dtype, x.device)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        hub = ctx.hub
        if g is None:
            if not hub.pending:
                return None, None
            shape, dtype, device = ctx.meta
            g = torch.zeros(shape, dtype=dtype, device=device)
        elif hub.pending and not g.is_contiguous():
            g = g.contiguous()
        for idx, rows in hub.pending:
            g.view(-1, g.shape[-1]).index_add_(0, idx, rows.to(g.dtype))
        hub.pending.clear()
        return g, None


class _SelectRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flat_idx, hub):
        ctx.hub = hub
        ctx.save_for_backward(flat_idx)
        return x.reshape(-1, x.shape[-1])[flat_idx]

    @staticmethod
    def backward(ctx, g):
        (flat_idx,) = ctx.saved_tensors
        ctx.hub.pending.append((flat_idx, g))
        return None, None, None


def grad_hub(x, hub):
    """Identity on `x`; row-sparse gradients registered on `hub` are folded into x's gradient in place."""
    return _HubFn.apply(x, hub) if x.requires_grad else x


def select_rows(x, flat_idx, hub):
    """x.reshape(-1, C)[flat_idx] whose backward is a row-sparse update routed through `hub` (x must be the output
    of grad_hub(..., hub))."""
    if not x.requires_grad:
        return x.reshape(-1, x.shape[-1])[flat_idx]
    return _SelectRowsFn.apply(x, flat_idx, hub)


class _EmbedRowsFn(torch.autograd.Function):
    """weight[idx] for a SMALL table hit by many indices (the denoising class embedding: 11 rows, thousands of
    lookups, models/utils/ops.py:243).  The library backward sorts the indices and serialises the colliding rows
    (161 us at TAM-TR shapes); as a one-hot GEMM [rows, n]^T @ grad it is a few microseconds and deterministic."""

    @staticmethod
    def forward(ctx, weight, idx):
        ctx.save_for_backward(idx)
        ctx.n = weight.shape[0]
        ctx.wdt = weight.dtype
        return weight[idx]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        with torch.autocast("cuda", enabled=False):
            onehot = torch.zeros(idx.numel(), ctx.n, dtype=torch.float32, device=g.device)
            onehot.scatter_(1, idx.view(-1, 1), 1.0)
            gw = onehot.t() @ g.reshape(idx.numel(), -1).float()
        return gw.to(ctx.wdt), None


def embed_rows(weight, idx):
    if not (weight.is_cuda and weight.requires_grad and torch.is_grad_enabled()):
        return weight[idx]
    return _EmbedRowsFn.apply(weight, idx)


# ------------------------------------------------------------------------------------ token-major input projection
def _col_reduce2_partials(a, b, tok0, ntok):
    """Per-CTA partial sums [n_cta, 2, d] fp32 of (a, a*b) over tokens [tok0, tok0+ntok) of every image; the fold to
    fp64 totals happens inside tamtr_bn_*_coeffs."""
    B, Lv, d = a.shape
    n_cta = _lib.lib().tamtr_col_reduce2_ctas(B, ntok)
    partial = torch.empty(n_cta, 2, d, dtype=torch.float32, device=a.device)
    with _with_device(a):
        rc = _lib.lib().tamtr_col_reduce2(a.data_ptr(), b.data_ptr(), partial.data_ptr(), _lib.dtype_code(a), B, Lv, d,
                                          tok0, ntok, _lib.stream_ptr(a.device))
    _lib.check(rc, "col_reduce2")
    return partial


def _affine_rows(a, b, A, Bc, Cc, starts):
    B, Lv, d = a.shape
    out = torch.empty_like(a)
    st = (ctypes.c_int32 * len(starts))(*starts)
    with _with_device(a):
        rc = _lib.lib().tamtr_affine_rows(out.data_ptr(), a.data_ptr(), b.data_ptr() if b is not None else None,
                                          A.data_ptr(), Bc.data_ptr() if Bc is not None else None, Cc.data_ptr(),
                                          _lib.dtype_code(a), B, Lv, d, len(starts), st, _lib.stream_ptr(a.device))
    _lib.check(rc, "affine_rows")
    return out


class _InputProjFn(torch.autograd.Function):
    """input_proj of the heads (head.py:1202-1218: per level Conv2d(1x1, bias=False) + BatchNorm2d, then
    flatten(2).permute(0,2,1) and cat) computed directly in token-major layout:

      pre[b, s_l:e_l, :] = x_l[b]^T W_l^T            one batched GEMM per level, written in place into its slice
      train: mu, var = column statistics of the slice (tamtr_col_reduce2), running stats updated like nn.BatchNorm2d
      feats = pre * (gamma*rstd) + (beta - mu*gamma*rstd)                     (tamtr_affine_rows)

    backward:  d_beta = sum G, d_gamma = rstd * (sum G*pre - mu * sum G)       (tamtr_col_reduce2 on (G, pre))
               d_pre  = gamma*rstd * (G - d_beta/M - xhat*d_gamma/M)  as  A*G + Bc*pre + Cc  (tamtr_affine_rows)
               dW_l = sum_b x_l[b] d_pre_l[b],  dx_l[b] = W_l^T d_pre_l[b]^T   (library GEMMs)
    """

    @staticmethod
    def forward(ctx, bns, training, n_levels, *t):
        xs, ws, gammas, betas = t[:n_levels], t[n_levels:2 * n_levels], t[2 * n_levels:3 * n_levels], t[3 * n_levels:]
        lp = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else xs[0].dtype
        if lp not in (torch.float32, torch.bfloat16):
            lp = torch.float32
        B, d = xs[0].shape[0], ws[0].shape[0]
        hw = [x.shape[2] * x.shape[3] for x in xs]
        starts = [sum(hw[:i]) for i in range(n_levels)]
        Lv = sum(hw)
        dev = xs[0].device
        pre = torch.empty(B, Lv, d, dtype=lp, device=dev)
        xl, wl = [], []
        for l in range(n_levels):
            a = xs[l].to(lp).flatten(2)                                   # [B, C, HW]  (NCHW, no copy when already lp)
            w = ws[l].reshape(d, -1).to(lp)                               # [d, C]
            torch.bmm(a.transpose(1, 2), w.t().unsqueeze(0).expand(B, -1, -1), out=pre[:, starts[l]:starts[l] + hw[l]])
            xl.append(a)
            wl.append(w)
        scale = torch.empty(n_levels, d, dtype=torch.float32, device=dev)
        shift = torch.empty_like(scale)
        stats = torch.empty(n_levels, 2, d, dtype=torch.float64, device=dev)          # (mu, rstd) per level
        for l, bn in enumerate(bns):
            batch_stats = training or bn.running_mean is None
            partial = _col_reduce2_partials(pre, pre, starts[l], hw[l]) if batch_stats else None
            update = batch_stats and training and bn.running_mean is not None
            mom = 0.0
            if update:
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            rm = rv = None
            if bn.running_mean is not None and (update or not batch_stats):
                rm, rv = bn.running_mean, bn.running_var
                if rm.dtype != torch.float32 or rv.dtype != torch.float32 or not rm.is_contiguous():
                    rm, rv = rm.float().contiguous(), rv.float().contiguous()
            g32, b32 = gammas[l].detach().float().contiguous(), betas[l].detach().float().contiguous()
            with _with_device(pre):
                rc = _lib.lib().tamtr_bn_forward_coeffs(
                    None if partial is None else partial.data_ptr(), 0 if partial is None else partial.shape[0],
                    float(B * hw[l]), g32.data_ptr(), b32.data_ptr(), float(bn.eps), int(batch_stats), float(mom),
                    None if rm is None else rm.data_ptr(), None if rv is None else rv.data_ptr(), scale[l].data_ptr(),
                    shift[l].data_ptr(), stats[l, 0].data_ptr(), stats[l, 1].data_ptr(), d, _lib.stream_ptr(dev))
            _lib.check(rc, "bn_forward_coeffs")
            if update:
                with torch.no_grad():
                    if rm is not bn.running_mean:
                        bn.running_mean.copy_(rm)
                        bn.running_var.copy_(rv)
                    bn.num_batches_tracked += 1
        mus = [stats[l, 0] for l in range(n_levels)]
        rstds = [stats[l, 1] for l in range(n_levels)]
        feats = _affine_rows(pre, None, scale, None, shift, starts)
        ctx.save_for_backward(pre, *xl, *wl, *gammas, *mus, *rstds)
        ctx.meta = (n_levels, hw, starts, training, [x.dtype for x in xs], [w.dtype for w in ws],
                    [g.dtype for g in gammas], [x.shape for x in xs], [w.shape for w in ws])
        return feats

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, G):
        n, hw, starts, training, xdt, wdt, gdt, xshape, wshape = ctx.meta
        sv = ctx.saved_tensors
        pre = sv[0]
        xl, wl, gammas = sv[1:1 + n], sv[1 + n:1 + 2 * n], sv[1 + 2 * n:1 + 3 * n]
        mus, rstds = sv[1 + 3 * n:1 + 4 * n], sv[1 + 4 * n:1 + 5 * n]
        G = G.contiguous().to(pre.dtype)
        B, Lv, d = pre.shape
        dev = pre.device
        A = torch.empty(n, d, dtype=torch.float32, device=dev)
        Bc = torch.empty_like(A)
        Cc = torch.empty_like(A)
        dgb = torch.empty(n, 2, d, dtype=torch.float32, device=dev)
        d_gamma, d_beta = [], []
        for l in range(n):
            partial = _col_reduce2_partials(G, pre, starts[l], hw[l])
            g32 = gammas[l].detach().float().contiguous()
            with _with_device(pre):
                rc = _lib.lib().tamtr_bn_backward_coeffs(
                    partial.data_ptr(), partial.shape[0], float(B * hw[l]), g32.data_ptr(), mus[l].data_ptr(),
                    rstds[l].data_ptr(), int(training), A[l].data_ptr(), Bc[l].data_ptr(), Cc[l].data_ptr(),
                    dgb[l, 0].data_ptr(), dgb[l, 1].data_ptr(), d, _lib.stream_ptr(dev))
            _lib.check(rc, "bn_backward_coeffs")
            d_gamma.append(dgb[l, 0].to(gdt[l]))
            d_beta.append(dgb[l, 1].to(gdt[l]))
        dpre = _affine_rows(G, pre if training else None, A, Bc if training else None, Cc, starts)
        d_x, d_w = [], []
        for l in range(n):
            dp = dpre[:, starts[l]:starts[l] + hw[l]]                                    # [B, HW, d]
            if ctx.needs_input_grad[3 + n + l]:
                d_w.append(torch.bmm(xl[l], dp).float().sum(0).t().reshape(wshape[l]).to(wdt[l]))
            else:
                d_w.append(None)
            if ctx.needs_input_grad[3 + l]:
                gx = torch.bmm(wl[l].t().unsqueeze(0).expand(B, -1, -1), dp.transpose(1, 2))   # [B, C, HW]
                d_x.append(gx.reshape(xshape[l]).to(xdt[l]))
            else:
                d_x.append(None)
        return (None, None, None, *d_x, *d_w, *d_gamma, *d_beta)


def input_proj_tokens(xs, projs, training):
    """xs: list of NCHW maps; projs: ModuleList of Sequential(Conv2d(1x1, bias=False), BatchNorm2d).
    Returns feats [B, sum(H_l*W_l), d] and the level shapes (head.py:1202-1218)."""
    _lib.require_cuda(*xs)
    convs, bns = [p[0] for p in projs], [p[1] for p in projs]
    n = len(xs)
    feats = _InputProjFn.apply(bns, training, n, *xs, *[c.weight for c in convs], *[b.weight for b in bns],
                               *[b.bias for b in bns])
    return feats, [[x.shape[2], x.shape[3]] for x in xs]


TOPK_KERNEL = os.environ.get("TAMTR_TOPK", "1") != "0"      # 0: the library's torch.topk (A/B switch)


def topk_rows(scores, k):
    """`torch.topk(scores, k, dim=1).indices` for the query selection (head.py:1240, :437): scores [B, n] -> int64 [B, k],
    best first.  fp32 CUDA rows that fit in shared memory go through tamtr_topk_rows (one launch; equal scores come out
    in index order, where torch leaves the order unspecified); anything else is the library call."""
    if not (TOPK_KERNEL and scores.is_cuda and scores.dtype == torch.float32 and scores.dim() == 2):
        return torch.topk(scores, k, dim=1).indices
    B, n = scores.shape
    L = _lib.lib()
    if L.tamtr_topk_rows_supported(n, k) != 2:           # rows too long for shared memory: the library is faster
        return torch.topk(scores, k, dim=1).indices
    scores = scores.detach().contiguous()
    out = torch.empty(B, k, dtype=torch.int64, device=scores.device)
    with _with_device(scores):
        _lib.check(L.tamtr_topk_rows(scores.data_ptr(), out.data_ptr(), None, B, n, k, _lib.stream_ptr(scores.device)),
                   "topk_rows")
    return out


def rank_tokens(feats, valid_u8, enc_linear, enc_norm, score_linear):
    """Query-selection ranking (head.py:1229-1237) without autograd: max over classes of
    enc_score_head(LayerNorm(enc_output.0(valid * feats))) for every token -> [B, Lv] fp32."""
    B, Lv, d = feats.shape
    # every tensor handed to the kernel is built with autocast OFF: under an outer autocast a plain `@` would silently
    # produce bf16 where the C ABI expects fp32
    with torch.no_grad(), torch.autocast("cuda", enabled=False):
        lp = feats.dtype
        f2 = feats.reshape(B * Lv, d)
        Wp = score_linear.weight.float() * enc_norm.weight.float()                 # [nc, d]
        nc = Wp.shape[0]
        npad = (nc + 15) // 16 * 16            # keeps the skinny GEMM on 16-byte aligned (tensor-core) kernels
        Wpad = torch.zeros(npad, d, dtype=torch.float32, device=feats.device)
        Wpad[:nc] = Wp
        if lp == torch.float32:
            E = f2 @ enc_linear.weight.float().t()
            raw = E @ Wpad.t()
        else:
            E = f2 @ enc_linear.weight.to(lp).t()
            raw = torch.mm(E, Wpad.to(lp).t(), out_dtype=torch.float32)
        eb = enc_linear.bias.float().contiguous()
        bw = (Wp @ eb).float().contiguous()
        sw = Wp.sum(1).float().contiguous()
        ck = (score_linear.weight.float() @ enc_norm.bias.float() + score_linear.bias.float()).float().contiguous()
        out = torch.empty(B, Lv, dtype=torch.float32, device=feats.device)
        raw = raw.contiguous()
        assert raw.dtype == torch.float32 and bw.dtype == torch.float32 and ck.dtype == torch.float32
        with _with_device(feats):
            rc = _lib.lib().tamtr_rank_tokens(E.data_ptr(), raw.data_ptr(), eb.data_ptr(), valid_u8.data_ptr(),
                                              bw.data_ptr(), sw.data_ptr(), ck.data_ptr(), out.data_ptr(),
                                              _lib.dtype_code(E), B, Lv, d, nc, npad, float(enc_norm.eps),
                                              _lib.stream_ptr(feats.device))
        _lib.check(rc, "rank_tokens")
    return out
