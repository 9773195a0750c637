"""Detection loss of the TAM-TR / RT-DETR heads, matching included, ON THE DEVICE and with fixed shapes.

Mirrors (same class names, constructor arguments, forward signatures, loss-dict keys):
  HungarianMatcher       ultralytics/models/utils/ops.py:12-121
  DETRLoss               ultralytics/models/utils/loss.py:14-373
  RTDETRDetectionLoss    ultralytics/models/utils/loss.py:376-443
The arithmetic of VarifocalLoss / FocalLoss (ultralytics/utils/loss.py:135-178) and bbox_iou's IoU / RIOU branches
(ultralytics/utils/metrics.py:71-130) lives in csrc/detloss.cu.

The reference builds the cost matrix on the GPU, copies it to the host and runs scipy.optimize.linear_sum_assignment
per image (ops.py:116-117) -- once per decoder layer plus once for the encoder proposals, i.e. four host round trips
in the middle of every training step -- and evaluates the losses layer by layer in ~150 small eager ops.  Here a
batch's ground truth lives in padded device tensors (DeviceTargets); `tamtr_match_cost` prices every query against its
own image's boxes, `tamtr_linear_sum_assignment_padded` (csrc/assign.cu: the same shortest-augmenting-path algorithm,
fp64 arithmetic and tie order as SciPy, one warp per image) assigns all layers in one launch, and
`tamtr_detection_loss` evaluates the losses of all layers AND their gradients in one pass.  Nothing depends on the
host knowing the ground-truth counts: no synchronisation, one captured CUDA graph serves every batch.
"""
import torch
import torch.nn as nn

from . import _lib

__all__ = ("HungarianMatcher", "DETRLoss", "RTDETRDetectionLoss", "DeviceTargets", "linear_sum_assignment", "match_padded")


# ------------------------------------------------------------------------------------------------ batch bookkeeping
class _Groups:
    """Everything about a batch that depends only on its per-image ground-truth counts (host-known): segment starts for
    the assignment kernel, the image index of every matched pair.  Built once per distinct (gt_groups, nq, device)."""
    _cache = {}

    def __init__(self, gt_groups, nq, device):
        self.counts = [min(nq, int(n)) for n in gt_groups]
        self.n_pairs = sum(self.counts)
        starts = [0]
        for n in gt_groups:
            starts.append(starts[-1] + int(n))
        outs = [0]
        for c in self.counts[:-1]:
            outs.append(outs[-1] + c)
        self.gt_start = torch.tensor(starts, dtype=torch.int32).to(device)
        self.out_start = torch.tensor(outs, dtype=torch.int32).to(device)
        self.pair_image = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(self.counts)]
                                    or [torch.zeros(0, dtype=torch.long)]).to(device)
        self.max_gt = max([int(n) for n in gt_groups] + [0])
        self.total_gt = starts[-1]
        # (image, slot) of every ground truth in a padded [B, G, .] layout, and the counts on the device
        self.gt_image = torch.cat([torch.full((int(n),), i, dtype=torch.long) for i, n in enumerate(gt_groups)]
                                  or [torch.zeros(0, dtype=torch.long)]).to(device)
        self.gt_slot = torch.cat([torch.arange(int(n), dtype=torch.long) for n in gt_groups]
                                 or [torch.zeros(0, dtype=torch.long)]).to(device)
        self.counts_dev = torch.tensor([int(n) for n in gt_groups], dtype=torch.int32).to(device)
        # global gt index of column g of image b's own (padded) cost matrix; padding columns repeat a valid index
        pad = torch.zeros(len(gt_groups), max(self.max_gt, 1), dtype=torch.long)
        for b, n in enumerate(gt_groups):
            if n:
                pad[b, :n] = torch.arange(starts[b], starts[b] + int(n))
                pad[b, n:] = starts[b]
        self.pad_index = pad.to(device)

    @classmethod
    def get(cls, gt_groups, nq, device):
        key = (tuple(int(n) for n in gt_groups), int(nq), str(device))
        hit = cls._cache.get(key)
        if hit is None:
            if len(cls._cache) > 64:
                cls._cache.clear()
            hit = cls._cache[key] = cls(gt_groups, nq, device)
        return hit


def linear_sum_assignment(C, gt_groups, padded=False):
    """C: [n_layers, bs, nq, total_gt] fp32 cost matrices on the device, image b owns the next gt_groups[b] columns
    (padded=True: [n_layers, bs, nq, >= max_gt], image b owns the first gt_groups[b] columns of its own matrix).
    Returns (image idx [P], query idx [n_layers, P], global gt idx [n_layers, P]) with P = sum_b min(nq, gt_groups[b]);
    per image the pairs are in ascending query order, exactly what scipy.optimize.linear_sum_assignment returns for
    C[l, b][:, columns of b] (ops.py:116-121)."""
    _lib.require_cuda(C)
    n_layers, bs, nq, c_cols = C.shape
    grp = _Groups.get(gt_groups, nq, C.device)
    assert len(gt_groups) == bs and (c_cols >= grp.max_gt if padded else grp.total_gt == c_cols)
    out_q = torch.empty(n_layers, grp.n_pairs, dtype=torch.long, device=C.device)
    out_g = torch.empty_like(out_q)
    if grp.n_pairs:
        C = C.contiguous().float()
        with torch.cuda.device(C.device):
            rc = _lib.lib().tamtr_linear_sum_assignment(C.data_ptr(), grp.gt_start.data_ptr(), grp.out_start.data_ptr(),
                                                        out_q.data_ptr(), out_g.data_ptr(), n_layers, bs, nq, c_cols,
                                                        int(padded), grp.max_gt, grp.n_pairs, _lib.stream_ptr(C.device))
        _lib.check(rc, "linear_sum_assignment")
    return grp.pair_image, out_q, out_g


class DeviceTargets:
    """The ground truth of a batch in FIXED-shape device tensors -- boxes [B, G, 4] (cx, cy, w, h), cls [B, G] int64,
    count [B] int32 -- so that one captured CUDA graph (denoising group, matching, loss) serves every batch: `load()`
    refreshes the tensors in place; everything that depends on the counts is decided inside the kernels.

    `dn_capacity`: slots reserved for denoising queries (the bucket's Lq = dn_capacity + num_queries).  The reference
    sizes its group per batch, 2 * max_gt * max(1, num_dn // max_gt) (ops.py:194-195, 242: at most 2 * num_dn while no
    image has more than num_dn boxes, 2 * max_gt beyond); `capacity_for()` returns the bucket that holds it."""

    def __init__(self, bs, max_gt, device, dn_capacity=None, num_dn=100):
        self.bs, self.max_gt, self.device = int(bs), int(max_gt), torch.device(device)
        self.boxes = torch.zeros(bs, max_gt, 4, dtype=torch.float32, device=device)
        self.cls = torch.zeros(bs, max_gt, dtype=torch.int64, device=device)
        self.count = torch.zeros(bs, dtype=torch.int32, device=device)
        self.dn_capacity, self.num_dn = dn_capacity, int(num_dn)
        pin = self.device.type == "cuda"
        self._h_boxes = torch.zeros(bs, max_gt, 4, dtype=torch.float32, pin_memory=pin)
        self._h_cls = torch.zeros(bs, max_gt, dtype=torch.int64, pin_memory=pin)
        self._h_count = torch.zeros(bs, dtype=torch.int32, pin_memory=pin)
        self._copied = None         # event after the last asynchronous copy out of the pinned staging buffers

    @staticmethod
    def dn_queries(max_gt, num_dn=100):
        return 0 if max_gt <= 0 or num_dn <= 0 else 2 * max_gt * max(1, num_dn // max_gt)

    @classmethod
    def capacity_for(cls, max_gt, num_dn=100):
        """Bucket capacities: 2 * num_dn (every batch without an image of more than num_dn boxes: 200 by default), then
        growing by ~1.5x in multiples of 64 (320, 512, 768, 1152 ...)."""
        need, cap = cls.dn_queries(max_gt, num_dn), max(2 * int(num_dn), 1)
        while cap < need:
            cap = (cap * 3 // 2 + 63) // 64 * 64
        return cap

    def to(self, device):
        if torch.device(device) == self.device:
            return self
        out = DeviceTargets(self.bs, self.max_gt, device, self.dn_capacity, self.num_dn)
        out.boxes.copy_(self.boxes)
        out.cls.copy_(self.cls)
        out.count.copy_(self.count)
        return out

    def load(self, batch, non_blocking=True):
        """batch: the reference's dict -- 'cls' [n], 'bboxes' [n, 4], 'gt_groups' [B ints], ground truths grouped image
        by image (data/dataset collate order) -- on the host or the device, or another DeviceTargets."""
        if isinstance(batch, DeviceTargets):
            self.boxes.copy_(batch.boxes, non_blocking=non_blocking)
            self.cls.copy_(batch.cls, non_blocking=non_blocking)
            self.count.copy_(batch.count, non_blocking=non_blocking)
            return self
        groups = [int(n) for n in batch["gt_groups"]]
        if len(groups) != self.bs or max(groups + [0]) > self.max_gt:
            raise RuntimeError(f"tamtr_b200: batch with gt_groups {groups} does not fit DeviceTargets(bs={self.bs}, "
                               f"max_gt={self.max_gt})")
        if self.dn_capacity is not None and self.dn_queries(max(groups + [0]), self.num_dn) > self.dn_capacity:
            raise RuntimeError("tamtr_b200: this batch needs more denoising slots than the captured bucket holds")
        boxes, cls = batch["bboxes"], batch["cls"].reshape(-1)
        if boxes.is_cuda:                       # already on the device (the reference's loss call): one scatter each
            grp = _Groups.get(groups, 1, boxes.device)
            self.boxes.zero_()
            self.cls.zero_()
            if grp.total_gt:
                self.boxes[grp.gt_image, grp.gt_slot] = boxes.float()
                self.cls[grp.gt_image, grp.gt_slot] = cls.long()
            self.count.copy_(grp.counts_dev)
            return self
        if self._copied is not None:
            self._copied.synchronize()          # the previous load's copies must have left the staging buffers
        self._h_boxes.zero_()
        self._h_cls.zero_()
        start = 0
        for b, n in enumerate(groups):
            self._h_boxes[b, :n] = boxes[start:start + n]
            self._h_cls[b, :n] = cls[start:start + n]
            start += n
        self._h_count.copy_(torch.tensor(groups, dtype=torch.int32))
        self.boxes.copy_(self._h_boxes, non_blocking=non_blocking)
        self.cls.copy_(self._h_cls, non_blocking=non_blocking)
        self.count.copy_(self._h_count, non_blocking=non_blocking)
        if self.device.type == "cuda":
            self._copied = self._copied or torch.cuda.Event()
            self._copied.record(torch.cuda.current_stream(self.device))
        return self

    @classmethod
    def from_batch(cls, batch, device, max_gt=None, num_dn=100):
        groups = [int(n) for n in batch["gt_groups"]]
        g = max(groups + [1]) if max_gt is None else max_gt
        return cls(len(groups), g, device, cls.capacity_for(max(groups + [0]), num_dn), num_dn).load(batch)


def _as_targets(batch, device):
    return batch if isinstance(batch, DeviceTargets) else DeviceTargets.from_batch(batch, device)


def match_padded(pred_bboxes, pred_scores, tgt, alpha=0.25, gamma=2.0, gains=(2.0, 5.0, 2.0)):
    """pred_* [NL, B, Q, .] -> match [NL, B, Q] int32 (ground-truth index inside the image, or -1): the cost matrices of
    ops.py:77-112 (tamtr_match_cost) and scipy's assignment per (layer, image) (tamtr_linear_sum_assignment_padded), two
    launches, no host synchronisation."""
    _lib.require_cuda(pred_bboxes, pred_scores)
    NL, B, Q, nc = pred_scores.shape
    G = tgt.max_gt
    pb = pred_bboxes.detach().float().contiguous()
    ps = pred_scores.detach().float().contiguous()
    cost = torch.empty(NL, B, Q, G, dtype=torch.float32, device=pb.device)
    match = torch.empty(NL, B, Q, dtype=torch.int32, device=pb.device)
    lib = _lib.lib()
    with torch.cuda.device(pb.device):
        st = _lib.stream_ptr(pb.device)
        _lib.check(lib.tamtr_match_cost(pb.data_ptr(), ps.data_ptr(), tgt.boxes.data_ptr(), tgt.cls.data_ptr(), cost.data_ptr(),
                                        NL, B, Q, G, nc, alpha, gamma, gains[0], gains[1], gains[2], st), "match_cost")
        _lib.check(lib.tamtr_linear_sum_assignment_padded(cost.data_ptr(), tgt.count.data_ptr(), match.data_ptr(), NL, B, Q,
                                                          G, st), "linear_sum_assignment_padded")
    return match


class _DetLossFn(torch.autograd.Function):
    """(class, bbox, giou) losses of NL layers [NL, 3], forward values and prediction gradients from ONE kernel pass
    (tamtr_detection_loss); the backward only scales the stored derivatives by the upstream gradients."""

    @staticmethod
    def forward(ctx, pred_bboxes, pred_scores, tgt, match, dn_group, num_dn, use_vfl, gains):
        NL, B, Q, nc = pred_scores.shape
        pb, ps = pred_bboxes.float().contiguous(), pred_scores.float().contiguous()
        dev = pb.device
        partial = torch.empty(NL, B, 3, dtype=torch.float32, device=dev)
        out = torch.empty(NL * 3 + 1, dtype=torch.float32, device=dev)
        d_l1, d_giou = torch.empty_like(pb), torch.empty_like(pb)
        d_cls = torch.empty_like(ps)
        with torch.cuda.device(dev):
            rc = _lib.lib().tamtr_detection_loss(
                pb.data_ptr(), ps.data_ptr(), tgt.boxes.data_ptr(), tgt.cls.data_ptr(), tgt.count.data_ptr(),
                None if match is None else match.data_ptr(), partial.data_ptr(), out.data_ptr(), d_l1.data_ptr(),
                d_giou.data_ptr(), d_cls.data_ptr(), NL, B, Q, tgt.max_gt, nc, int(dn_group), int(num_dn), int(use_vfl),
                gains[0], gains[1], gains[2], _lib.stream_ptr(dev))
        _lib.check(rc, "detection_loss")
        ctx.save_for_backward(d_l1, d_giou, d_cls, out)
        ctx.gains = gains
        ctx.dtypes = (pred_bboxes.dtype, pred_scores.dtype)
        return out[:NL * 3].view(NL, 3)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        d_l1, d_giou, d_cls, out = ctx.saved_tensors
        NL = g.shape[0]
        scale = (g.float() * out[NL * 3]).view(NL, 3, 1, 1, 1)                # upstream gradient / matched pairs
        gb = (scale[:, 1] * ctx.gains[1]) * d_l1 + (scale[:, 2] * ctx.gains[2]) * d_giou
        gs = (scale[:, 0] * ctx.gains[0]) * d_cls
        return gb.to(ctx.dtypes[0]), gs.to(ctx.dtypes[1]), None, None, None, None, None, None


def _loss_dict(per_layer, aux, postfix, with_class=True):
    """loss.py:328-373: the last layer's three losses + the sums over the auxiliary layers."""
    out = {}
    if with_class:
        out[f"loss_class{postfix}"] = per_layer[-1, 0]
    out[f"loss_bbox{postfix}"] = per_layer[-1, 1]
    out[f"loss_giou{postfix}"] = per_layer[-1, 2]
    if aux:
        rest = per_layer[:-1].sum(0)
        out[f"loss_class_aux{postfix}"] = rest[0]
        out[f"loss_bbox_aux{postfix}"] = rest[1]
        out[f"loss_giou_aux{postfix}"] = rest[2]
    return out


class HungarianMatcher(nn.Module):
    """ops.py:12-121.  forward() keeps the reference's signature and return value (a list of (query idx, gt idx) per
    image); match_layers() / match_padded() are the batched forms the loss uses (all decoder layers in one launch)."""

    def __init__(self, cost_gain=None, use_fl=True, with_mask=False, num_sample_points=12544, alpha=0.25, gamma=2.0):
        super().__init__()
        if cost_gain is None:
            cost_gain = {'class': 1, 'bbox': 5, 'giou': 2, 'mask': 1, 'dice': 1}
        if with_mask:
            raise NotImplementedError("tamtr_b200: mask costs are commented out in the reference (ops.py:124-150)")
        self.cost_gain = cost_gain
        self.use_fl = use_fl
        self.with_mask = with_mask
        self.num_sample_points = num_sample_points
        self.alpha = alpha
        self.gamma = gamma

    def match_padded(self, pred_bboxes, pred_scores, tgt):
        if not self.use_fl:
            raise NotImplementedError("tamtr_b200: the softmax class cost (use_fl=False) is not on TAM-TR's path")
        g = self.cost_gain
        return match_padded(pred_bboxes, pred_scores, tgt, float(self.alpha), float(self.gamma),
                            (float(g['class']), float(g['bbox']), float(g['giou'])))

    def match_layers(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups):
        """pred_* [n_layers, bs, nq, .] -> (image idx [P], query idx [n_layers, P], global gt idx [n_layers, P]), the pairs
        of each image in ascending query order -- what scipy returns per image (ops.py:116-121).  (This list form needs the
        number of pairs on the host; the loss itself works on the fixed-shape match array.)"""
        n_l = pred_bboxes.shape[0]
        dev = pred_bboxes.device
        if sum(gt_groups) == 0:
            z = torch.zeros(0, dtype=torch.long, device=dev)
            e = torch.zeros(n_l, 0, dtype=torch.long, device=dev)
            return z, e, e.clone()
        tgt = DeviceTargets.from_batch({"bboxes": gt_bboxes, "cls": gt_cls, "gt_groups": gt_groups}, dev)
        match = self.match_padded(pred_bboxes, pred_scores, tgt)                # [n_l, bs, nq]
        grp = _Groups.get(gt_groups, pred_scores.shape[2], dev)
        hit = match >= 0
        img, q = hit[0].nonzero(as_tuple=True)                                  # the same count in every layer
        qs = torch.stack([hit[l].nonzero(as_tuple=True)[1] for l in range(n_l)])
        lay = torch.arange(n_l, device=dev).unsqueeze(1)
        gs = match[lay, img.unsqueeze(0), qs].long() + grp.gt_start[:-1].long()[img].unsqueeze(0)
        return img, qs, gs

    def forward(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, masks=None, gt_mask=None):
        bs, nq, nc = pred_scores.shape
        if sum(gt_groups) == 0:
            return [(torch.tensor([], dtype=torch.long), torch.tensor([], dtype=torch.long)) for _ in range(bs)]
        _, q, g = self.match_layers(pred_bboxes.unsqueeze(0), pred_scores.unsqueeze(0), gt_bboxes, gt_cls, gt_groups)
        counts = _Groups.get(gt_groups, nq, pred_scores.device).counts
        return list(zip(q[0].split(counts), g[0].split(counts)))


class DETRLoss(nn.Module):
    """loss.py:14-373 (varifocal / focal classification loss, L1 + RIoU box losses, auxiliary losses per layer) on the
    kernels of csrc/detloss.cu: matching and the losses of ALL layers in four launches, fixed shapes, no host round trip
    (the reference matches and evaluates layer by layer, loss.py:232-243, 293-300, with scipy on the host)."""

    def __init__(self, nc=80, loss_gain=None, aux_loss=True, use_fl=True, use_vfl=False, use_sl=False, use_emasl=False,
                 use_svfl=False, use_emasvfl=False, use_uni_match=False, uni_match_ind=0):
        super().__init__()
        if use_sl or use_emasl or use_svfl or use_emasvfl:
            raise NotImplementedError("tamtr_b200: slide-loss variants are not on TAM-TR's path (nn/tasks.py:578)")
        if loss_gain is None:
            loss_gain = {'class': 1, 'bbox': 5, 'giou': 2, 'no_object': 0.1, 'mask': 1, 'dice': 1}
        self.nc = nc
        self.matcher = HungarianMatcher(cost_gain={'class': 2, 'bbox': 5, 'giou': 2})
        self.loss_gain = loss_gain
        self.aux_loss = aux_loss
        self.fl = use_fl            # (the reference stores FocalLoss() / VarifocalLoss() modules or None here; only their
        self.vfl = use_vfl          #  truth value is used: the arithmetic of utils/loss.py:135-178 lives in the kernel)
        self.use_uni_match = use_uni_match
        self.uni_match_ind = uni_match_ind
        self.device = None

    def _get_loss_layers(self, pred_bboxes, pred_scores, tgt, dn_group=False, num_dn=100, postfix=''):
        """pred_* [n_l, bs, nq, .] -> the loss dict of loss.py:328-373 for these layers."""
        if not self.fl:
            raise NotImplementedError("tamtr_b200: plain BCE classification (use_fl=False) is not on TAM-TR's path")
        if self.use_uni_match:
            raise NotImplementedError("tamtr_b200: use_uni_match is not on TAM-TR's path (nn/tasks.py:578)")
        if not self.aux_loss:
            pred_bboxes, pred_scores = pred_bboxes[-1:], pred_scores[-1:]
        match = None if dn_group else self.matcher.match_padded(pred_bboxes, pred_scores, tgt)
        g = self.loss_gain
        per_layer = _DetLossFn.apply(pred_bboxes, pred_scores, tgt, match, dn_group, num_dn, bool(self.vfl),
                                     (float(g['class']), float(g['bbox']), float(g['giou'])))
        return _loss_dict(per_layer, self.aux_loss, postfix)

    def forward(self, pred_bboxes, pred_scores, batch, postfix='', **kwargs):
        """pred_bboxes [l, b, query, 4], pred_scores [l, b, query, nc]; batch: the reference's dict (cls / bboxes /
        gt_groups) or a DeviceTargets.  kwargs: dn_group / num_dn select the denoising targets (RTDETRDetectionLoss)."""
        _lib.require_cuda(pred_bboxes)
        self.device = pred_bboxes.device
        if pred_scores is None:
            raise NotImplementedError("tamtr_b200: box-only losses (pred_scores=None) are not on TAM-TR's path")
        tgt = _as_targets(batch, pred_bboxes.device)
        return DETRLoss._get_loss_layers(self, pred_bboxes, pred_scores, tgt, kwargs.get('dn_group', False),
                                         kwargs.get('num_dn', 100), postfix)


class RTDETRDetectionLoss(DETRLoss):
    """loss.py:376-443: the detection loss plus the denoising loss on the CDN queries (fixed matches: the positive
    copies of the denoising layout, get_dn_match_indices)."""

    def forward(self, preds, batch, dn_bboxes=None, dn_scores=None, dn_meta=None):
        # (explicit DETRLoss.forward instead of super(): patch.enable() binds this function onto the reference's class)
        pred_bboxes, pred_scores = preds
        tgt = _as_targets(batch, pred_bboxes.device)
        total_loss = DETRLoss.forward(self, pred_bboxes, pred_scores, tgt)
        if dn_meta is not None:
            # the kernel re-derives the layout from the counts: num_dn > 0 = the head's `num_denoising` (device-built
            # groups record it); the reference's dn_meta only carries the resulting number of groups (passed negated)
            num_dn = dn_meta.get('num_dn_cfg') or -int(dn_meta['dn_num_group'])
            total_loss.update(DETRLoss.forward(self, dn_bboxes, dn_scores, tgt, postfix='_dn', dn_group=True, num_dn=num_dn))
        else:
            total_loss.update({f'{k}_dn': torch.zeros((), device=self.device) for k in list(total_loss.keys())})
        return total_loss
