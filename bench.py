"""Headline benchmark: TAM-TR detection-head (MEH) forward + backward, images/s, on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): the head TAMTR.yaml:67 builds -- ManbaWorldDecoder(nc=10, ch=[128,256,512],
hd=512, nq=100, ndp=4, nh=8, ndl=3) = 3 deformable decoder layers over a 160^2/80^2/40^2 pyramid (33 600 tokens,
d=512, 8 heads x 64) + the text-guided contrastive classification branch on 10 synthetic 512-d text embeddings --
train mode with a contrastive-denoising group (20..100 synthetic VisDrone-like boxes per image), bf16 autocast,
batch 16 per GPU, forward + backward (+ one NCCL all-reduce of the flat gradient buffer when N > 1).  VSSBlocks are
identity (their CUDA extension is not part of the reference tree).  Weights random-init, data
synthetic.

One JSON line on stdout (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` is the same step
through the public module API with the step's inputs coming from pinned HOST memory (copy stream, double-buffered)
and the loss read back to the host every step.  `roofline` is for the dominant hand-written kernel (the sampler's
backward); per-kernel device times come from CUDA events the C library records on the launching stream around its
own launches during an instrumented pass over the same K steps.  `cpu_baseline` / `--impl reference`: the CPU
restatement of the reference's own PyTorch path (oracle/head_ref.py, fp32, all host threads) -- /root/reference is
pure Python and does not exist on the GPU box, so the "port" stands in for it (kind = "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

_JSON_OUT = sys.stdout      # main() replaces it by a private duplicate of fd 1

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NC, CH, HD, NQ, NDP, NH, NDL = 10, (128, 256, 512), 512, 100, 4, 8, 3
SIZES = (160, 80, 40)
BATCH_PER_GPU = 16
METRIC = "head fwd+bwd images/sec @640^2"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst: the kernel is timed alone)"
    return 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"


def gate_conv_roofline(dev, iters=20):
    """Kernel 2 (the tensor-bound one, BASELINE.json configs[2] at its largest point: 6400 image tokens x C=256, 10 text
    tokens, batch 16): the fused 3x3 projection + BatchNorm affine + text gate on tcgen05.  Device time per launch from
    the library's own CUDA events on the launching stream; three rotating buffer sets (315 MB > L2)."""
    from tamtr_b200 import _lib, ops
    B, C, nh, H, W, N = 16, 256, 8, 80, 80, 10
    g = torch.Generator().manual_seed(7)
    xs = [torch.randn(B, C, H, W, generator=g).bfloat16().to(dev).contiguous(memory_format=torch.channels_last)
          for _ in range(3)]
    w = (torch.randn(C, C, 3, 3, generator=g) * (2.0 / (9 * C)) ** 0.5).bfloat16().to(dev)
    s = (1.0 + 0.1 * torch.randn(C, generator=g)).to(dev)
    t = (0.1 * torch.randn(C, generator=g)).to(dev)
    guide = (0.3 * torch.randn(B, N, nh, C // nh, generator=g)).to(dev)
    bias = torch.zeros(nh, device=dev)
    with torch.no_grad():
        gates = [ops.max_sigmoid_gate(x, guide, bias, nh) for x in xs]
        for i in range(5):
            ops.gate_conv3x3(xs[i % 3], w, s, t, gates[i % 3], nh)
        torch.cuda.synchronize(dev)
        _lib.profile_enable(True)
        for i in range(iters):
            ops.gate_conv3x3(xs[i % 3], w, s, t, gates[i % 3], nh)
        torch.cuda.synchronize(dev)
        ms, n = _lib.profile_read()["gate_conv3x3_tc_fwd"]
        _lib.profile_enable(False)
    flops = 2.0 * B * H * W * 9 * C * C
    us = ms / n * 1e3
    peak, src = tensor_peak()
    return {"bound": "tensor", "kernel": "gate_conv3x3_pair_kernel (tcgen05.mma cta_group::2, TMEM, TMA)",
            "achieved": flops / us / 1e6, "peak": peak, "unit": "TFLOP/s", "frac": flops / us / 1e6 / peak,
            "traffic": None, "peak_source": src, "algorithmic_flops": flops, "avg_us": us, "launches": n,
            "workload": "BTA-PAN text-guided 3x3 projection + BN + gate, B=16 C=256 80x80 (6400 tokens) N=10, bf16"}


def infer_config5(dev, iters=20):
    """Eval-mode forward of the MEH head at BASELINE.json configs[4] (1280x1280: levels 320^2/160^2/80^2, 900 queries,
    batch 1 per GPU-iteration), bf16 autocast, replayed as one CUDA graph (dp.HeadInferStep); VSSBlocks identity / on."""
    from tamtr_b200 import dp
    from tamtr_b200.head import ManbaWorldDecoder
    out = {"unit": "images/s per GPU", "workload": "MEH head eval forward, 1280x1280 (134 400 tokens), 900 queries, batch 1, "
                                                   "bf16, CUDA graph; inputs resident"}
    g = torch.Generator().manual_seed(99)
    xs = [torch.randn(1, c, s, s, generator=g).bfloat16().to(dev) for c, s in zip(CH, (320, 160, 80))]
    text = torch.nn.functional.normalize(torch.randn(1, NC, 512, generator=g), dim=-1).to(dev)
    for key, vss in (("vss_identity", False), ("vss_on", True)):
        torch.manual_seed(1234)
        m = ManbaWorldDecoder(NC, list(CH), HD, 900, NDP, NH, NDL, vss=vss).to(dev).eval()
        step = dp.HeadInferStep(m, (xs, text), autocast=torch.bfloat16)
        for _ in range(3):
            step.run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            step.run()
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        out[key] = {"ms_per_image": ms, "value": 1e3 / ms, "launches_of_ours": step.launches_per_step}
        del step, m
        torch.cuda.empty_cache()
    return out


def synthetic_targets(seed, B, lo=20, hi=100):
    """VisDrone-shaped ground truth: n ~ U{20..100} small boxes per image, 10 classes."""
    g = torch.Generator().manual_seed(seed)
    groups = [int(torch.randint(lo, hi + 1, (1,), generator=g)) for _ in range(B)]
    groups[0] = hi                    # pin the largest group so every rank/step has the same query count
    n = sum(groups)
    boxes = torch.cat([torch.rand(n, 2, generator=g), 0.01 + 0.29 * torch.rand(n, 2, generator=g)], -1)
    cls = torch.randint(0, NC, (n,), generator=g)
    idx = torch.cat([torch.full((k,), i, dtype=torch.long) for i, k in enumerate(groups)])
    return {"cls": cls, "bboxes": boxes, "batch_idx": idx, "gt_groups": groups}


def bind_to_gpu_numa_node(local_rank):
    """Multi-GPU runs: pin this process to the CPUs NVML reports as local to its GPU BEFORE the pinned host buffers are
    allocated, so that first touch places them on the GPU's NUMA node.  (With 8 ranks streaming 184 MB per step each from
    wherever the allocator put them, the host side of the H2D copies -- not the PCIe links -- bounded the end-to-end number.)
    Best effort: any failure leaves the affinity as it was."""
    try:
        import pynvml
        pynvml.nvmlInit()
        index = local_rank
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        if vis:
            entries = [e.strip() for e in vis.split(",") if e.strip()]
            if local_rank < len(entries):
                index = int(entries[local_rank]) if entries[local_rank].isdigit() else None
        handle = (pynvml.nvmlDeviceGetHandleByIndex(index) if index is not None
                  else pynvml.nvmlDeviceGetHandleByUUID(entries[local_rank]))
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def synthetic_inputs(seed, B, dtype):
    g = torch.Generator().manual_seed(seed)
    xs = [torch.randn(B, c, s, s, generator=g).to(dtype) for c, s in zip(CH, SIZES)]
    text = torch.nn.functional.normalize(torch.randn(B, NC, 512, generator=g), dim=-1)
    return xs, text


def surrogate_loss_fn(out):
    """Cheap scalar every head output feeds (--loss surrogate; the round-1 numbers before the device-side matcher)."""
    db, ds, eb, es = out[:4]
    return (db.float().square().mean() + 0.1 * ds.float().sigmoid().mean()
            + eb.float().square().mean() + 0.1 * es.float().sigmoid().mean())


def split_outputs(out):
    """ultralytics/nn/tasks.py:606-621: split the denoising queries off and prepend the encoder proposals."""
    dec_bboxes, dec_scores, enc_bboxes, enc_scores, dn_meta = out
    dn_bboxes = dn_scores = None
    if dn_meta is not None:
        dn_bboxes, dec_bboxes = torch.split(dec_bboxes, dn_meta["dn_num_split"], dim=2)
        dn_scores, dec_scores = torch.split(dec_scores, dn_meta["dn_num_split"], dim=2)
    dec_bboxes = torch.cat([enc_bboxes.unsqueeze(0), dec_bboxes])
    dec_scores = torch.cat([enc_scores.unsqueeze(0), dec_scores])
    return dec_bboxes, dec_scores, dn_bboxes, dn_scores, dn_meta


def make_detection_loss(batch, dev):
    """The reference's training loss (nn/tasks.py:578-624: RTDETRDetectionLoss(use_vfl=True) on the encoder proposals +
    every decoder layer + the denoising queries) with the Hungarian matching on the device."""
    from tamtr_b200.loss import RTDETRDetectionLoss
    crit = RTDETRDetectionLoss(nc=NC, use_vfl=True)
    targets = {"cls": batch["cls"].to(dev), "bboxes": batch["bboxes"].to(dev), "gt_groups": batch["gt_groups"]}

    def fn(out):
        db, ds, dnb, dns, meta = split_outputs(out)
        f = (lambda t: None if t is None else t.float())
        return sum(crit((db.float(), ds.float()), targets, dn_bboxes=f(dnb), dn_scores=f(dns), dn_meta=meta).values())
    return fn


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        mx = max((int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------- bytes
def sampler_bytes(B, Lq, value_bytes=2, L=3, P=4):
    """ALGORITHMIC bytes of one sampler launch (DESIGN.md 'Roofline accounting').
    fwd: each value byte the gather needs, once (min(dense slab, gathered)) + fp32 loc/attn + output.
    bwd: grad_out + the same value bytes + loc/attn read + grad_loc/grad_attn write + the dense grad_value written
         once (its zero-fill is a separate memset node and is NOT counted here, nor timed in this kernel)."""
    Lv = sum(s * s for s in SIZES)
    d, Dh = HD, HD // NH
    gathered = B * Lq * NH * L * P * 4 * Dh
    val = min(B * Lv * d, gathered) * value_bytes
    locw = B * Lq * NH * L * P * 3 * 4
    out = B * Lq * d * value_bytes
    fwd = val + locw + out
    bwd = out + val + 2 * locw + min(B * Lv * d, gathered) * value_bytes
    return fwd, bwd


# ----------------------------------------------------------------------------------------------------- CPU port
def cpu_reference_step(sample_images, threads, steps, warmup, seed=1234, loss_kind="surrogate"):
    """The reference's own PyTorch path on the host cores (oracle port): fp32, train mode, fwd + bwd."""
    from oracle import head_ref, loss_ref       # the ONLY things on this path: no product code, no CUDA
    torch.set_num_threads(threads)
    torch.manual_seed(seed)
    sd = head_ref.meh_state_dict(NC, CH, HD, NDL, NH, seed=seed)
    sd = {k: (v.requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    xs, text = synthetic_inputs(seed, sample_images, torch.float32)
    batch = synthetic_targets(seed, sample_images)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cdn = head_ref.cdn_group(batch, NC, NQ, sd["denoising_class_embed.weight"])
        out = head_ref.head(sd, "", xs, NQ, NDL, NH, training=True, text=text, cdn=cdn)
        if loss_kind == "surrogate":
            loss = head_ref.surrogate_loss(*out)
        else:       # the reference's RTDETRDetectionLoss with scipy's linear_sum_assignment on the host (ops.py:117)
            meta = head_ref.cdn_meta(batch, NQ) if cdn is not None else None
            db, ds, dnb, dns, meta = split_outputs((*out, meta))
            loss = sum(loss_ref.rtdetr_detection_loss(db, ds, batch["bboxes"], batch["cls"], batch["gt_groups"], NC,
                                                      dn_bboxes=dnb, dn_scores=dns, dn_meta=meta).values())
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sample_images * len(times) / sum(times), sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = 2
    steps, warmup = min(args.steps, 8), min(args.warmup, 1)
    ips, sec = cpu_reference_step(sample, cores, steps, warmup, loss_kind=args.loss)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.loss),
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} images/step of the same workload (S-yaml head, train fwd+bwd, fp32), "
                                   f"{steps} steps after {warmup} warm-up, torch CPU with {cores} threads"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), file=_JSON_OUT, flush=True)


LOSS_NAMES = {"detection": "RTDETRDetectionLoss(use_vfl) on encoder proposals + 3 decoder layers + denoising queries, "
                           "Hungarian matching per layer (nn/tasks.py:578-624)",
              "surrogate": "mean-square / mean-sigmoid scalar over all outputs"}


def config_dict(loss_kind="surrogate", vss=False):
    return {"loss": LOSS_NAMES[loss_kind], "vss_blocks": "selective-scan kernels" if vss else "identity",
            "workload": "TAM-TR MEH head (ManbaWorldDecoder nc=10 ch=[128,256,512] hd=512 nq=100 ndl=3) "
                        "+ text-guided cls branch, train fwd+bwd, CDN 20..100 gt/img, pyramid 160^2/80^2/40^2 @640^2",
            "batch_per_gpu": BATCH_PER_GPU, "text_tokens": NC, "text_dim": 512,
            "l2": "inputs larger than L2 (per-step working set > 1 GB vs 126 MB L2); no explicit flush"}


# ----------------------------------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="tamtr_b200", choices=["tamtr_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--loss", default="surrogate", choices=["detection", "surrogate"],
                    help="what drives the backward.  surrogate (default): a cheap scalar over all head outputs, so that the "
                         "timed region is the head's forward + backward (BASELINE.json's metric); detection: the "
                         "reference's RTDETRDetectionLoss on top (Hungarian matching on the device; scipy on the host in "
                         "the reference arm).  With the default, the detection-loss step is also timed and reported as "
                         "`with_detection_loss`.")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--vss", action="store_true",
                    help="run the head's three VSSBlocks for real (selective-scan kernels) instead of identities.  Off by "
                         "default: the reference cannot run them without its un-vendored CUDA extension, so the reference "
                         "arm and the head-level parity fixtures are identity-VSS; with the default the VSS-on step is "
                         "also timed once and reported as `with_vss`.")
    ap.add_argument("--launch-list", action="store_true",
                    help="eager steps only (no e2e / instrumented pass / CPU baseline): the command to run under "
                         "`ncu --metrics gpu__time_duration.sum` for profiles/launches_*.csv")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: keep a private handle on it and point fd 1 at stderr for everything else
    # (NCCL prints its version banner to stdout at init, whatever NCCL_DEBUG_FILE says)
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    import tamtr_b200
    from tamtr_b200 import _lib, dp
    from tamtr_b200.head import ManbaWorldDecoder

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    ws = int(os.environ.get("WORLD_SIZE", 1))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for this path)"
    numa_cpus = bind_to_gpu_numa_node(local) if ws > 1 else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if ws > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()

    B = BATCH_PER_GPU
    torch.manual_seed(1234)                                # same initial weights on every rank (DDP broadcast equivalent)
    model = ManbaWorldDecoder(NC, list(CH), HD, NQ, NDP, NH, NDL, vss=args.vss).to(dev).train()
    # two synthetic batches per rank in pinned host memory (bf16 activations, as a bf16 neck would hand them over)
    host = []
    for j in range(2):
        xs, text = synthetic_inputs(1234 + rank * 7 + j, B, torch.bfloat16)
        host.append(([x.pin_memory() for x in xs], text.pin_memory()))
    batch = synthetic_targets(1234 + rank, B)
    plan = model.plan_cdn(batch)
    Lq = plan.n_dn + NQ

    loss_fn = make_detection_loss(batch, dev) if args.loss == "detection" else surrogate_loss_fn
    step = dp.HeadTrainStep(model, loss_fn, (host[0][0], host[0][1], plan), autocast=torch.bfloat16,
                            use_graph=not (args.no_graph or args.launch_list))
    if args.launch_list:
        for _ in range(args.warmup + args.steps):
            step.run()
        torch.cuda.synchronize(dev)
        print(json.dumps({"launch_list": True, "steps": args.steps, "warmup": args.warmup}), file=_JSON_OUT, flush=True)
        return
    h2d = sum(x.numel() * x.element_size() for x in host[0][0]) + host[0][1].numel() * host[0][1].element_size()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ device-resident timing ("value")
    for _ in range(args.warmup):
        step.run()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _lib.launch_count()
    e0.record()
    for _ in range(args.steps):
        step.run()
    e1.record()
    barrier()
    sec = dp.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)
    clk = clocks.stop() if rank == 0 else None
    launches = (step.launches_per_step * args.steps) if step.graph is not None else (_lib.launch_count() - launches0)
    value = ws * B * args.steps / sec

    # ------------------------------------------------------------------ end to end from pinned host memory ("e2e")
    copy_stream = torch.cuda.Stream(dev)
    staging = [([torch.empty_like(x, device=dev) for x in host[0][0]], torch.empty_like(host[0][1], device=dev))
               for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            for d, h in zip(staging[s][0], host[s][0]):
                d.copy_(h, non_blocking=True)
            staging[s][1].copy_(host[s][1], non_blocking=True)
            ready[s].record(copy_stream)

    def e2e_loop(n):
        """Every step: H2D of its inputs (copy stream, overlapping the previous step), the step, D2H of its loss.  The host
        reads step i's loss right after it has enqueued step i+1 (a training loop logging its loss does the same), so the
        device is never idle waiting for the host; every step's loss is read."""
        cur = torch.cuda.current_stream(dev)
        for s in range(2):
            consumed[s].record(cur)
        stage(0)
        losses = []
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                stage(i + 1)                       # next step's H2D overlaps this step's compute
            cur.wait_event(ready[s])
            step.load_inputs((staging[s][0], staging[s][1], None))   # device->static-buffer copy (graph inputs)
            consumed[s].record(cur)
            loss = step.run()
            loss_host[s].copy_(loss, non_blocking=True)
            loss_done[s].record(cur)
            if i > 0:
                loss_done[1 - s].synchronize()
                losses.append(float(loss_host[1 - s]))
        if n > 0:
            loss_done[(n - 1) % 2].synchronize()
            losses.append(float(loss_host[(n - 1) % 2]))
        assert len(losses) == n
        return losses

    e2e_loop(args.warmup)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    barrier()
    e2e_sec = dp.max_over_ranks(t0.elapsed_time(t1) / 1e3, dev)
    e2e_value = ws * B * args.steps / e2e_sec
    # the host->device link on its own (explains e2e when it, not the step, is the longer leg)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(copy_stream):
        c0.record(copy_stream)
        for _ in range(4):
            for d, h in zip(staging[0][0], host[0][0]):
                d.copy_(h, non_blocking=True)
        c1.record(copy_stream)
    copy_stream.synchronize()
    h2d_gbs = 4 * sum(x.numel() * x.element_size() for x in host[0][0]) / (c0.elapsed_time(c1) / 1e3) / 1e9

    # ------------------------------------------------------------------ per-kernel device times (instrumented pass)
    kern = {}
    if rank == 0:
        _lib.profile_enable(True)
        graph, step.graph = step.graph, None       # eager pass over the same buffers: the library brackets each of
        for _ in range(args.steps):                # its launches with CUDA events on the launching stream
            step.run(reduce=False)
        torch.cuda.synchronize(dev)
        step.graph = graph
        kern = _lib.profile_read()
        _lib.profile_enable(False)
    barrier()

    if rank == 0:
        peak, peak_src = peaks()
        fwd_b, bwd_b = sampler_bytes(B, Lq)
        per_kernel = {k: {"avg_us": ms / n * 1e3, "launches_per_step": n / args.steps} for k, (ms, n) in kern.items()}
        bwd_us = per_kernel.get("msda_bwd", {}).get("avg_us")
        fwd_us = per_kernel.get("msda_fwd", {}).get("avg_us")
        roof = {"bound": "hbm", "kernel": "msda_bwd_kernel<bf16,LPC=8,NS=12>", "achieved": bwd_b / (bwd_us * 1e3) if bwd_us else None,
                "peak": peak, "unit": "GB/s", "frac": (bwd_b / (bwd_us * 1e3) / peak) if bwd_us else None,
                "traffic": None, "peak_source": peak_src, "algorithmic_bytes": bwd_b, "avg_us": bwd_us,
                "fwd": {"kernel": "msda_fwd_kernel<bf16,LPC=8,NS=12>", "achieved": fwd_b / (fwd_us * 1e3) if fwd_us else None,
                        "frac": (fwd_b / (fwd_us * 1e3) / peak) if fwd_us else None, "algorithmic_bytes": fwd_b,
                        "avg_us": fwd_us}}
        ncu = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(ncu):
            try:
                roof["traffic"] = json.load(open(ncu)).get("msda_bwd_dram_bytes_per_launch")
            except Exception:
                pass
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": ws, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(config_dict(args.loss, args.vss), queries=Lq, cuda_graph=step.graph is not None,
                               parallelism=f"dp{ws}" if ws > 1 else "single", cpus_bound_to_gpu_numa_node=numa_cpus),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": e2e_sec / args.steps * 1e3, "h2d_link_gbs_measured": h2d_gbs,
                        "h2d_ms_per_step_at_link_rate": h2d / h2d_gbs / 1e6},
                "gpu_launches": int(launches),
                "roofline": roof, "kernels": per_kernel}
        if ws == 1 and args.loss == "surrogate":
            try:        # the same step with the reference's detection loss on top (device-side Hungarian matching)
                step2 = dp.HeadTrainStep(model, make_detection_loss(batch, dev), (host[0][0], host[0][1], plan),
                                         autocast=torch.bfloat16, use_graph=not args.no_graph)
                for _ in range(args.warmup):
                    step2.run()
                torch.cuda.synchronize(dev)
                d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                d0.record()
                for _ in range(args.steps):
                    step2.run()
                d1.record()
                torch.cuda.synchronize(dev)
                dsec = d0.elapsed_time(d1) / 1e3
                line["with_detection_loss"] = {"value": B * args.steps / dsec, "unit": "images/s",
                                               "ms_per_step": dsec / args.steps * 1e3, "loss": LOSS_NAMES["detection"],
                                               "loss_value": float(step2.loss)}
                del step2
            except Exception as e:
                line["with_detection_loss"] = {"error": str(e)[:200]}
        if ws == 1 and not args.vss:
            try:        # the same step with the three VSSBlocks running (head.py:1092-1098,1134)
                torch.manual_seed(1234)
                model_v = ManbaWorldDecoder(NC, list(CH), HD, NQ, NDP, NH, NDL, vss=True).to(dev).train()
                for blk in model_v.VSSBlocks:
                    blk.drop_path.drop_prob = 0.0            # stochastic depth draws random numbers: not graph-replayable
                step3 = dp.HeadTrainStep(model_v, surrogate_loss_fn, (host[0][0], host[0][1], plan),
                                         autocast=torch.bfloat16, use_graph=not args.no_graph, warmup=1)
                nv = max(2, args.steps // 4)
                step3.run()
                torch.cuda.synchronize(dev)
                v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                v0.record()
                for _ in range(nv):
                    step3.run()
                v1.record()
                torch.cuda.synchronize(dev)
                vsec = v0.elapsed_time(v1) / 1e3
                line["with_vss"] = {"value": B * nv / vsec, "unit": "images/s", "ms_per_step": vsec / nv * 1e3, "steps": nv,
                                    "note": "VSSBlocks on the selective-scan kernels (fp32 scan as vmamba.py:985 forces), "
                                            "drop_path 0; parity for the scan: published recurrence (oracle) + forward cross-check against "
                                            "vLLM's mamba_ssm kernel; the reference's own extension is not in its tree"}
                del step3, model_v
                torch.cuda.empty_cache()
            except Exception as e:
                line["with_vss"] = {"error": str(e)[:200]}
        if ws == 1:
            try:        # BASELINE.json configs[4]: inference at 1280x1280, 900 queries, one image per GPU-iteration
                line["infer_1280"] = infer_config5(dev)
            except Exception as e:
                line["infer_1280"] = {"error": str(e)[:200]}
        if ws == 1:
            try:
                line["roofline_tensor"] = gate_conv_roofline(dev)
                ncu = os.path.join(ROOT, "profiles", "traffic.json")
                if os.path.exists(ncu):
                    line["roofline_tensor"]["traffic"] = json.load(open(ncu)).get("gate_conv_dram_bytes_per_launch")
            except Exception as e:      # the headline line must not depend on the secondary kernel
                line["roofline_tensor"] = {"error": str(e)[:200]}
        if not args.no_cpu_baseline and ws == 1:
            cores = os.cpu_count() or 1
            ips, s = cpu_reference_step(2, cores, 2, 1, loss_kind=args.loss)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "2 images/step of the same workload (fp32 train fwd+bwd), 2 timed steps "
                                              f"after 1 warm-up, torch CPU {cores} threads, {s:.1f} s/step"}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
