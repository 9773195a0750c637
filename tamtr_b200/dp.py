"""Data-parallel head training step (forward + backward + gradient all-reduce), one process per GPU.

The reference trains with torch DDP: batch sharded by DistributedSampler, one bucketed NCCL all-reduce (mean) of all
gradients per backward (ultralytics/engine/trainer.py:241, 252; data/build.py:104).  The path shards by image with
no data-path collective; the only exchange step is that gradient all-reduce.  B200-first restatement:

* the gradients are packed into ONE flat fp32 buffer by a single multi-tensor copy at the end of the backward -> the
  exchange is a single NCCL all-reduce over NVLink/NVSwitch, issued right after the backward;
* forward + backward of the step are captured ONCE into a CUDA graph (static input buffers, the denoising group is
  planned on the host per batch exactly as the reference does and only its embedding gather is in the graph), so the
  ~1.5k small launches of the step cost one graph launch instead of Python/launch latency;
* host inputs arrive through pinned staging buffers on a copy stream (double-buffered), overlapping the previous
  step's compute.

`HeadTrainStep` is device-agnostic where it can be: with `use_graph=False` and a CPU module it runs the same
flat-gradient / all-reduce logic under the gloo backend, which is how tests/test_dp_cpu.py covers the N>1 path.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, rank, world_size):
    """Contiguous, balanced shard of `n_items` independent units (images) for `rank` -- no collective involved."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class FlatGrads:
    """One flat buffer for all parameter gradients (DDP's gradient_as_bucket_view with a single bucket).

    Gradients are NOT accumulated into the buffer by autograd (that is one tiny `grad += g` launch per parameter,
    ~130 for the MEH head): the backward writes fresh .grad tensors, `gather()` packs them with one multi-tensor
    copy, and the exchange is a single all-reduce over the flat buffer."""

    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def clear(self):
        for p in self.params:
            p.grad = None

    def gather(self):
        """Pack the parameters' .grad tensors into the flat buffer (one fused multi-tensor copy)."""
        dst, src = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                dst.append(v)
                src.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, src)
        return self.flat

    def all_reduce_mean(self):
        """One all-reduce (sum) + scale: the DDP semantics of trainer.py:241 (mean over ranks)."""
        rank, ws = world()
        if ws > 1:
            if dist.get_backend() == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)      # the division happens inside the collective
            else:                                                     # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                self.flat.div_(ws)
        return self.flat

    def scatter(self):
        """Point every .grad at its (reduced) slice of the flat buffer, e.g. before an optimizer step."""
        for p, v in zip(self.params, self.views):
            p.grad = v


class _FusedParamCast(torch.autograd.Function):
    """All low-precision weight copies of a step with ONE multi-tensor copy (and one more for their gradients).

    Under autocast every nn.Linear casts its fp32 weight and bias to bf16 on use and autograd casts the gradients
    back: ~180 tiny launches per step for the MEH head.  The values are identical (same round-to-nearest cast)."""

    @staticmethod
    def forward(ctx, dtype, *params):
        outs = [torch.empty_like(p, dtype=dtype) for p in params]
        torch._foreach_copy_(outs, list(params))
        ctx.src = [(p.dtype, p.shape, p.device) for p in params]
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        outs = [None if g is None else torch.empty(shape, dtype=dt, device=dev)
                for g, (dt, shape, dev) in zip(grads, ctx.src)]
        dst = [o for o in outs if o is not None]
        if dst:
            torch._foreach_copy_(dst, [g for g in grads if g is not None])
        return (None, *outs)


def lowp_param_names(module):
    """Parameters that autocast would cast on every use: weights/biases of Linear layers and of MultiheadAttention."""
    names = []
    for mn, m in module.named_modules():
        if isinstance(m, torch.nn.Linear):
            names += [f"{mn}.weight"] + ([f"{mn}.bias"] if m.bias is not None else [])
        elif isinstance(m, torch.nn.MultiheadAttention) and m.in_proj_weight is not None:
            names += [f"{mn}.in_proj_weight"] + ([f"{mn}.in_proj_bias"] if m.in_proj_bias is not None else [])
    return [n.lstrip(".") for n in names]


class HeadTrainStep:
    """forward + backward (+ all-reduce) of a detection head on static buffers.

    module     : tamtr_b200.head.ManbaWorldDecoder / RTDETRDecoder (train mode), or any module for the CPU tests
    loss_fn    : maps the module's outputs to a scalar
    example    : tuple of example inputs (tensors / CdnPlan / None) fixing every shape
    autocast   : torch dtype or None
    use_graph  : capture forward+backward into a CUDA graph (CUDA only).  As for any whole-network capture, eager
                 forward/backward passes of the SAME module instance done earlier in the process must have run on a
                 side stream (autograd's gradient accumulators remember the stream of their first use); the warm-up
                 here does.
    """

    def __init__(self, module, loss_fn, example, autocast=None, use_graph=True, warmup=3, fused_param_cast=True):
        self.module, self.loss_fn, self.autocast = module, loss_fn, autocast
        self.cast_names = lowp_param_names(module) if (fused_param_cast and autocast is not None) else []
        self.flat = FlatGrads(module.parameters())
        self.pack_grads = world()[1] > 1 or not use_graph     # single process: gradients can stay where autograd put them
        self.device = self.flat.flat.device
        self.cuda = self.device.type == "cuda"
        self.use_graph = use_graph and self.cuda
        self.static = [self._to_static(a) for a in example]
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.graph = None
        self.launches_per_step = None
        if self.use_graph:
            self._capture(warmup)

    def _to_static(self, a):
        if isinstance(a, torch.Tensor):
            return a.detach().to(self.device).clone()
        if isinstance(a, (list, tuple)):
            return type(a)(self._to_static(x) for x in a)
        if hasattr(a, "to") and hasattr(a, "materialize"):      # CdnPlan
            return a.to(self.device)
        return a

    def _fwd_bwd(self):
        self.flat.clear()
        if self.autocast is not None:
            with torch.autocast(self.device.type, dtype=self.autocast):
                if self.cast_names:
                    named = dict(self.module.named_parameters())
                    lowp = _FusedParamCast.apply(self.autocast, *[named[n] for n in self.cast_names])
                    out = torch.func.functional_call(self.module, dict(zip(self.cast_names, lowp)), tuple(self.static))
                else:
                    out = self.module(*self.static)
        else:
            out = self.module(*self.static)
        loss = self.loss_fn(out)
        loss.backward()
        self.loss.copy_(loss.detach())
        if self.pack_grads:
            self.flat.gather()

    def _capture(self, warmup):
        from . import _lib
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._fwd_bwd()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self._fwd_bwd()
        self.launches_per_step = _lib.launch_count() - before
        torch.cuda.synchronize(self.device)

    def load_inputs(self, inputs, non_blocking=True):
        """Copy a new batch (same shapes) into the static buffers."""
        def cp(dst, src):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src, non_blocking=non_blocking)
            elif isinstance(dst, (list, tuple)):
                for d, s in zip(dst, src):
                    cp(d, s)
        for d, s in zip(self.static, inputs):
            cp(d, s)

    def run(self, reduce=True):
        """One step on whatever is in the static buffers; returns the (device) loss tensor."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._fwd_bwd()
        if reduce and self.pack_grads:
            self.flat.all_reduce_mean()
        return self.loss


class HeadInferStep:
    """Eval-mode forward of a detection head on static buffers, captured once into a CUDA graph.

    Inference is sharded by image, one process per GPU, no collective.  At the batch sizes that
    leaves per GPU the eager forward is launch-bound (the MEH head at 1280x1280 takes 5 ms for 1 image and for 8), so the
    whole forward -- input projection, query selection, the decoder layers with their sampler / projection / contrastive
    kernels -- is replayed as one graph.

    module   : tamtr_b200.head.ManbaWorldDecoder / RTDETRDecoder (put in eval mode here)
    example  : tuple of example inputs fixing every shape, e.g. (list of 3 feature maps, text)
    autocast : torch dtype or None
    run(inputs=None) copies `inputs` (same shapes) into the static buffers when given, replays, and returns the module's
    outputs -- static tensors that the next run() overwrites."""

    def __init__(self, module, example, autocast=None, use_graph=True, warmup=3):
        self.module, self.autocast = module.eval(), autocast
        self.device = next(module.parameters()).device
        self.static = [self._to_static(a) for a in example]
        self.out, self.graph, self.launches_per_step = None, None, None
        if use_graph and self.device.type == "cuda":
            from . import _lib
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                for _ in range(max(1, warmup)):
                    self._fwd()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.launch_count()
            with torch.cuda.graph(self.graph):
                self.out = self._fwd()
            self.launches_per_step = _lib.launch_count() - before
            torch.cuda.synchronize(self.device)

    def _to_static(self, a):
        if isinstance(a, torch.Tensor):
            return a.detach().to(self.device).clone()
        if isinstance(a, (list, tuple)):
            return type(a)(self._to_static(x) for x in a)
        return a

    def _fwd(self):
        with torch.no_grad():
            if self.autocast is not None:
                with torch.autocast(self.device.type, dtype=self.autocast):
                    return self.module(*self.static)
            return self.module(*self.static)

    def load_inputs(self, inputs, non_blocking=True):
        def cp(dst, src):
            if isinstance(dst, torch.Tensor):
                dst.copy_(src, non_blocking=non_blocking)
            elif isinstance(dst, (list, tuple)):
                for d, s_ in zip(dst, src):
                    cp(d, s_)
        for d, s_ in zip(self.static, inputs):
            cp(d, s_)

    def run(self, inputs=None):
        if inputs is not None:
            self.load_inputs(inputs)
        if self.graph is not None:
            self.graph.replay()
        else:
            self.out = self._fwd()
        return self.out


def max_over_ranks(seconds, device):
    """Multi-GPU timings are the max over ranks (never a wall clock of rank 0 alone)."""
    rank, ws = world()
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
