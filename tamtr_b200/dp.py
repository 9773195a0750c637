"""Data-parallel head training step (forward + backward + gradient all-reduce), one process per GPU.

The reference trains with torch DDP: batch sharded by DistributedSampler, one bucketed NCCL all-reduce (mean) of all
gradients per backward (ultralytics/engine/trainer.py:241, 252; data/build.py:104).  The path shards by image with
no data-path collective; the only exchange step is that gradient all-reduce.  B200-first restatement:

* the gradients live in ONE flat fp32 buffer cut into a few buckets in the order the backward finishes them (recorded
  during warm-up); each bucket is packed by one multi-tensor copy and all-reduced (NCCL AVG over NVLink/NVSwitch) on a
  communication stream AS SOON AS its last gradient exists, i.e. while the rest of the backward still runs -- from inside
  the captured graph (fork/join through events), so a replay carries its own overlapped exchange;
* parameters, gradients and the AdamW moments are flat fp32 buffers: the reference's optimizer step (clip_grad_norm_ 0.1
  + AdamW, engine/trainer.py:471-477) is three launches of csrc/optim.cu at the end of the same graph;
* forward + backward of the step are captured ONCE into a CUDA graph (static input buffers, the denoising group is
  planned on the host per batch exactly as the reference does and only its embedding gather is in the graph), so the
  ~1.5k small launches of the step cost one graph launch instead of Python/launch latency;
* host inputs arrive through pinned staging buffers on a copy stream (double-buffered), overlapping the previous
  step's compute.

`HeadTrainStep` is device-agnostic where it can be: with `use_graph=False` and a CPU module it runs the same
flat-gradient / all-reduce logic under the gloo backend, which is how tests/test_dp_cpu.py covers the N>1 path.
"""
import inspect

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

    def __init__(self, params, dtype=torch.float32, pad_to=4, sources=None):
        """`params`: in the order they are to be laid out; every parameter starts at a multiple of `pad_to` elements.
        `sources`: parameter -> the tensor whose .grad receives its gradient (default: the parameter itself; the
        training step back-propagates into low-precision leaf copies of the Linear weights)."""
        self.params = [p for p in params if p.requires_grad]
        self.src = [p if sources is None else sources.get(p, p) for p in self.params]
        dev = self.params[0].device
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + pad_to - 1) // pad_to * pad_to
        self.flat = torch.zeros(off, dtype=dtype, device=dev)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.buckets = [(0, len(self.params))]          # [first, last) parameter index of every bucket

    def set_buckets(self, n_buckets):
        """Cut the (already ordered) parameter list into `n_buckets` runs of about equal bytes."""
        total = self.flat.numel()
        cuts, first, acc = [], 0, 0
        for i, p in enumerate(self.params):
            acc += p.numel()
            if acc >= total * (len(cuts) + 1) / n_buckets and len(cuts) < n_buckets - 1 and i + 1 < len(self.params):
                cuts.append((first, i + 1))
                first = i + 1
        cuts.append((first, len(self.params)))
        self.buckets = cuts

    def bucket_slice(self, b):
        first, last = self.buckets[b]
        end = self.flat.numel() if last == len(self.params) else self.offsets[last]
        return self.flat[self.offsets[first]:end]

    def pack_bucket(self, b):
        first, last = self.buckets[b]
        groups = {}                                  # one multi-tensor copy per gradient dtype: a list that mixes bf16 and
        for t, v in zip(self.src[first:last], self.views[first:last]):      # fp32 sources makes _foreach_copy_ fall back
            if t.grad is None:                                               # to one copy kernel per tensor (~110 launches
                v.zero_()                                                    # per step for the MEH head)
            else:
                dst, src = groups.setdefault(t.grad.dtype, ([], []))
                dst.append(v)
                src.append(t.grad)
        for dst, src in groups.values():
            torch._foreach_copy_(dst, src)          # (converts low-precision gradients to the buffer's fp32 on the way)

    def clear(self):
        for t in self.src:
            t.grad = None

    def gather(self):
        """Pack all gradients into the flat buffer (one fused multi-tensor copy)."""
        whole, self.buckets = self.buckets, [(0, len(self.params))]
        self.pack_bucket(0)
        self.buckets = whole
        return self.flat

    def all_reduce_mean(self, tensor=None):
        """All-reduce (sum) + scale of the flat buffer or one bucket of it: the DDP semantics of trainer.py:241 (mean
        over ranks)."""
        t = self.flat if tensor is None else tensor
        rank, ws = world()
        if ws > 1:
            if dist.get_backend() == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.AVG)              # the division happens inside the collective
            else:                                                     # gloo (CPU tests) has no AVG
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                t.div_(ws)
        return t

    def scatter(self):
        """Point every .grad at its (reduced) slice of the flat buffer, e.g. before an optimizer step."""
        for p, v in zip(self.params, self.views):
            p.grad = v


def decays(module):
    """Parameter -> bool of the reference's optimizer groups (engine/trainer.py:654-662): weights get weight decay,
    anything with 'bias' in its name and the weights of normalisation layers do not."""
    norm = tuple(v for k, v in torch.nn.__dict__.items() if "Norm" in k and isinstance(v, type))
    out = {}
    for mname, m in module.named_modules():
        for pname, p in m.named_parameters(recurse=False):
            full = f"{mname}.{pname}" if mname else pname
            out[p] = not ("bias" in full or isinstance(m, norm))
    return out


class FlatAdamW:
    """clip_grad_norm_(max_norm) + AdamW (engine/trainer.py:471-477) over flat fp32 buffers.

    The parameters of `flat_grads` are MOVED into one flat buffer laid out like the gradient buffer (p.data becomes a
    view), so that the step is tamtr_adamw_flat: one pass for the global gradient norm, one pass that clips and updates
    parameters and both moments, a device-side step counter -- three launches, no host synchronisation (CUDA only)."""

    def __init__(self, flat_grads, decay_flags, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=0.1):
        """decay_flags: one bool per parameter of `flat_grads` (weight decay applies to it or not)."""
        from . import _lib
        self.g = flat_grads
        self.hyper = (float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), float(max_norm))
        dev = self.g.flat.device
        self.param = torch.zeros_like(self.g.flat)
        with torch.no_grad():
            for p, o in zip(self.g.params, self.g.offsets):
                v = self.param[o:o + p.numel()].view_as(p)
                v.copy_(p.data)
                p.data = v
        # one byte per group of four elements (every parameter starts at a multiple of four)
        self.decay4 = torch.zeros(self.param.numel() // 4, dtype=torch.uint8, device=dev)
        for p, o, d in zip(self.g.params, self.g.offsets, decay_flags):
            if d:
                self.decay4[o // 4:(o + p.numel() + 3) // 4] = 1
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.partial = torch.empty(_lib.lib().tamtr_optim_partials(self.param.numel()), dtype=torch.float32, device=dev)

    def step(self):
        from . import _lib
        lr, b1, b2, eps, wd, mx = self.hyper
        with torch.cuda.device(self.param.device):
            rc = _lib.lib().tamtr_adamw_flat(self.param.data_ptr(), self.g.flat.data_ptr(), self.exp_avg.data_ptr(),
                                            self.exp_avg_sq.data_ptr(), self.param.numel(), self.decay4.data_ptr(),
                                            self.partial.data_ptr(), self.step_count.data_ptr(), lr, b1, b2, eps, wd, mx,
                                            _lib.stream_ptr(self.param.device))
        _lib.check(rc, "adamw_flat")


class LowpLeaves:
    """Low-precision working copies of the parameters autocast would cast on every use, as LEAF tensors.

    Under autocast every nn.Linear casts its fp32 weight and bias to bf16 on use and autograd casts the gradients
    back: ~180 tiny launches per step for the MEH head.  Here the copies are refreshed by ONE multi-tensor copy at the
    start of a step (same round-to-nearest values) and the module is called on them; being leaves, each receives its
    gradient the moment the backward has finished it -- which is what lets a gradient bucket leave for its all-reduce
    while the rest of the backward is still running -- and FlatGrads converts it to fp32 while packing."""

    def __init__(self, module, names, dtype):
        named = dict(module.named_parameters())
        self.names = [n for n in names if named[n].requires_grad]
        self.params = [named[n] for n in self.names]
        self.leaves = [torch.empty_like(p, dtype=dtype).requires_grad_() for p in self.params]
        self.source_of = dict(zip(self.params, self.leaves))

    def refresh(self):
        with torch.no_grad():
            torch._foreach_copy_(self.leaves, [p.detach() for p in self.params])
        return dict(zip(self.names, self.leaves))


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
    """forward + backward (+ overlapped gradient exchange, + optimizer step) of a detection head on static buffers.

    module     : tamtr_b200.head.ManbaWorldDecoder / RTDETRDecoder (train mode), or any module for the CPU tests
    loss_fn    : maps the module's outputs -- or (outputs, the step's static inputs) -- to a scalar
    example    : tuple of example inputs (tensors / loss.DeviceTargets / CdnPlan / None) fixing every shape.  With a
                 DeviceTargets (ground truth padded to fixed shapes, denoising group built by a kernel from the counts)
                 the captured step serves EVERY batch: load_inputs() refreshes its tensors in place.  A host-planned
                 CdnPlan is baked into the capture (one batch only; kept for the reference's RNG-exact queries)
    autocast   : torch dtype or None
    use_graph  : capture the step into a CUDA graph (CUDA only).  As for any whole-network capture, eager
                 forward/backward passes of the SAME module instance done earlier in the process must have run on a
                 side stream (autograd's gradient accumulators remember the stream of their first use); the warm-up
                 here does.
    optimizer  : None, or a dict of FlatAdamW arguments (lr, betas, eps, weight_decay, max_norm): the reference's
                 optimizer_step at the end of every step (inside the graph)
    buckets    : world_size > 1 on CUDA: number of gradient buckets all-reduced from inside the step while the backward
                 is still running (0: one all-reduce of the whole buffer after the step, round 1's behaviour)
    """

    def __init__(self, module, loss_fn, example, autocast=None, use_graph=True, warmup=3, fused_param_cast=True,
                 optimizer=None, buckets=4, share=None):
        if share is not None:       # another capture (other input shapes) of the SAME training state: StepCache
            return self._init_shared(module, loss_fn, example, share, warmup)
        self.module, self.loss_fn, self.autocast = module, loss_fn, autocast
        # loss_fn(outputs) or loss_fn(outputs, static_inputs): the second form reads the step's own (in-place refreshed)
        # ground truth -- a loss.DeviceTargets among the inputs -- so that a captured step follows every new batch
        try:
            n_pos = sum(1 for q in inspect.signature(loss_fn).parameters.values()
                        if q.kind in (q.POSITIONAL_ONLY, q.POSITIONAL_OR_KEYWORD) and q.default is q.empty)
        except (TypeError, ValueError):
            n_pos = 1
        self._loss_takes_inputs = n_pos >= 2
        self.lowp = LowpLeaves(module, lowp_param_names(module), autocast) \
            if (fused_param_cast and autocast is not None) else None
        sources = {} if self.lowp is None else self.lowp.source_of
        params = [p for p in module.parameters() if p.requires_grad]
        self.device = params[0].device
        self.cuda = self.device.type == "cuda"
        self.use_graph = use_graph and self.cuda
        ws = world()[1]
        self.overlap = bool(buckets) and ws > 1 and self.cuda
        self.static = [self._to_static(a) for a in example]
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.flat, self.opt, self._live = None, None, False
        if self.overlap:
            params = self._completion_order(params, sources)
        self.flat = FlatGrads(params, sources=sources)
        if self.overlap:
            self.flat.set_buckets(int(buckets))
            self.comm = torch.cuda.Stream(self.device)
            self._bucket_of = {}
            for b, (first, last) in enumerate(self.flat.buckets):
                for t in self.flat.src[first:last]:
                    self._bucket_of[t] = b
            for t in self.flat.src:
                t.register_post_accumulate_grad_hook(self._on_grad)
        # a single process that keeps fp32 parameters all the way could leave the gradients where autograd put them
        self.pack_grads = ws > 1 or not self.use_graph or optimizer is not None or self.lowp is not None
        if optimizer is not None:
            if not self.cuda:
                raise RuntimeError("tamtr_b200: the flat AdamW step is a CUDA kernel (Not implemented on the CPU; is_cuda)")
            dec = decays(module)
            self.opt = FlatAdamW(self.flat, [dec.get(p, True) for p in self.flat.params], **optimizer)
        self.graph = None
        self.launches_per_step = None
        if self.use_graph:
            self._capture(warmup)

    def _init_shared(self, module, loss_fn, example, share, warmup):
        """Same module, low-precision leaves, flat gradient / parameter / moment buffers, communication stream and
        graph memory pool as `share`; only the static inputs and the captured graph are this step's own."""
        assert module is share.module
        for k in ("module", "autocast", "lowp", "device", "cuda", "use_graph", "overlap", "flat", "opt", "pack_grads",
                  "_loss_takes_inputs"):
            setattr(self, k, getattr(share, k))
        self.loss_fn = loss_fn
        self.static = [self._to_static(a) for a in example]
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self._live = False
        if self.overlap:
            self.comm, self._bucket_of = share.comm, share._bucket_of
            for t in self.flat.src:
                t.register_post_accumulate_grad_hook(self._on_grad)
        self.graph, self.launches_per_step = None, None
        if self.use_graph:
            self._capture(warmup, pool=share.graph.pool() if share.graph is not None else None)

    def _to_static(self, a):
        if isinstance(a, torch.Tensor):
            return a.detach().to(self.device).clone()
        if isinstance(a, (list, tuple)):
            return type(a)(self._to_static(x) for x in a)
        if hasattr(a, "to") and (hasattr(a, "materialize") or hasattr(a, "load")):      # CdnPlan / loss.DeviceTargets
            return a.to(self.device)
        return a

    # ---- gradient buckets -------------------------------------------------------------------------------------------
    def _completion_order(self, params, sources):
        """One eager backward on a side stream that records the order in which the parameters' gradients are finished:
        buckets cut along that order complete one after the other, the first long before the backward ends."""
        order, handles = [], []
        back = {sources.get(p, p): p for p in params}
        for t in back:
            handles.append(t.register_post_accumulate_grad_hook(lambda q: order.append(back[q])))
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self._forward_backward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for h in handles:
            h.remove()
        seen = set(order)
        for t in back:
            t.grad = None
        return order + [p for p in params if p not in seen]      # parameters without a gradient go last

    def _on_grad(self, t):
        if not self._live:
            return
        b = self._bucket_of[t]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch_bucket(b)

    def _launch_bucket(self, b):
        """Pack bucket b on the stream the backward is running on, then all-reduce it on the communication stream
        (fork through an event: under capture this becomes a parallel branch of the graph)."""
        self._launched.add(b)
        self.flat.pack_bucket(b)
        self.comm.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.comm):
            self.flat.all_reduce_mean(self.flat.bucket_slice(b))

    # ---- the step ---------------------------------------------------------------------------------------------------
    def _forward_backward(self):
        if self.autocast is not None:
            with torch.autocast(self.device.type, dtype=self.autocast):
                if self.lowp is not None:
                    out = torch.func.functional_call(self.module, self.lowp.refresh(), tuple(self.static))
                else:
                    out = self.module(*self.static)
        else:
            out = self.module(*self.static)
        loss = self.loss_fn(out, self.static) if self._loss_takes_inputs else self.loss_fn(out)
        loss.backward()
        self.loss.copy_(loss.detach())

    def _fwd_bwd(self):
        if self.flat is not None:
            self.flat.clear()
        if self.overlap:
            self._pending = [last - first for first, last in self.flat.buckets]
            self._launched, self._live = set(), True
        try:
            self._forward_backward()
        finally:
            self._live = False
        if self.overlap:
            for b in range(len(self.flat.buckets)):             # buckets holding a parameter that received no gradient
                if b not in self._launched:
                    self._launch_bucket(b)
            torch.cuda.current_stream(self.device).wait_stream(self.comm)      # join
        elif self.pack_grads:
            self.flat.gather()
        if self.opt is not None and (self.overlap or world()[1] == 1):
            self.opt.step()                                     # (otherwise after the all-reduce, in run())

    def _capture(self, warmup, pool=None):
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
        with torch.cuda.graph(self.graph, pool=pool):
            self._fwd_bwd()
        self.launches_per_step = _lib.launch_count() - before
        torch.cuda.synchronize(self.device)

    def load_inputs(self, inputs, non_blocking=True):
        """Copy a new batch (same shapes) into the static buffers.  A loss.DeviceTargets entry is refreshed in place
        (`load`: new boxes / classes / counts, any counts that fit its capacity).  Other non-tensor entries -- a host-planned
        CdnPlan carries one batch's denoising queries, mask and counts -- are baked into a captured graph and CANNOT be
        replaced: pass None to keep them, or build a new step (StepCache keeps one per denoising-capacity bucket)."""
        def cp(dst, src):
            if src is None or src is dst:
                return
            if isinstance(dst, torch.Tensor):
                dst.copy_(src, non_blocking=non_blocking)
            elif isinstance(dst, (list, tuple)):
                for d, s in zip(dst, src):
                    cp(d, s)
            else:
                if hasattr(dst, "load"):
                    dst.load(src, non_blocking=non_blocking)        # in-place update of a padded plan's device tensors
                else:
                    raise RuntimeError(f"tamtr_b200: static argument of type {type(dst).__name__} cannot be updated in a "
                                       "captured step; pass None to keep it or build a new HeadTrainStep")
        for d, s in zip(self.static, inputs):
            cp(d, s)

    def run(self, reduce=True):
        """One step on whatever is in the static buffers; returns the (device) loss tensor."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._fwd_bwd()
        if reduce and self.pack_grads and not self.overlap and world()[1] > 1:
            self.flat.all_reduce_mean()
            if self.opt is not None:
                self.opt.step()
        return self.loss


class StepCache:
    """Training steps for batches whose ground-truth counts vary: ONE captured HeadTrainStep per denoising-capacity bucket.

    The reference sizes its denoising group per batch -- Lq = nq + 2 * max_gt * max(1, num_dn // max_gt)
    (ultralytics/models/utils/ops.py:194-195, 242), 200 for max_gt <= 100 and 2 * max_gt above, data-dependent -- and
    rebuilds everything eagerly.  Here the ground truth travels in fixed-shape device tensors (loss.DeviceTargets) and the
    group is laid out by a kernel inside a bucket of `dn_capacity` slots (padding slots are masked out of the attention
    and ignored by the loss), so one graph serves every batch of its bucket: a VisDrone-like stream with <= 100 boxes per
    image never leaves the first bucket; a denser batch triggers ONE more capture (same weights, gradient / optimizer
    buffers and graph memory pool), after which that bucket replays too.

    make_inputs(batch, targets) -> the example tuple for HeadTrainStep with `targets` (a DeviceTargets) in it.
    run(batch, tensors) : batch = the reference's dict (cls / bboxes / gt_groups, host or device); tensors = the other
    inputs in the order of the example (None entries are kept)."""

    def __init__(self, module, loss_fn, make_inputs, num_dn=100, **step_args):
        self.module, self.loss_fn, self.make_inputs, self.num_dn = module, loss_fn, make_inputs, num_dn
        self.step_args = step_args
        self.steps = {}

    @staticmethod
    def bucket(max_gt, num_dn=100):
        """(dn_capacity, ground-truth slots per image) of the bucket a batch with this largest count falls into."""
        from .loss import DeviceTargets
        cap = DeviceTargets.capacity_for(max_gt, num_dn)
        return cap, (max(num_dn, 1) if cap <= 2 * num_dn else cap // 2)     # largest per-image count the bucket holds

    def step_for(self, batch):
        from .loss import DeviceTargets
        groups = [int(n) for n in batch["gt_groups"]]
        cap, slots = self.bucket(max(groups + [0]), self.num_dn)
        st = self.steps.get(cap)
        if st is None:
            dev = next(self.module.parameters()).device
            tgt = DeviceTargets(len(groups), slots, dev, cap, self.num_dn).load(batch)
            first = next(iter(self.steps.values()), None)
            args = dict(self.step_args) if first is None else {"warmup": self.step_args.get("warmup", 3)}
            st = self.steps[cap] = HeadTrainStep(self.module, self.loss_fn, self.make_inputs(batch, tgt), share=first, **args)
            st.targets = next(a for a in st.static if isinstance(a, DeviceTargets))
        return st

    def run(self, batch, tensors=None, reduce=True):
        st = self.step_for(batch)
        st.targets.load(batch)
        if tensors is not None:
            st.load_inputs([None if a is st.targets else t for a, t in zip(st.static, tensors)])
        return st.run(reduce=reduce)


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
