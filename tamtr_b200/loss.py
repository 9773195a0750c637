"""Detection loss of the TAM-TR / RT-DETR heads with the query <-> ground-truth matching ON THE DEVICE.

Mirrors (same class names, constructor arguments, forward signatures, loss-dict keys):
  HungarianMatcher       ultralytics/models/utils/ops.py:12-121
  DETRLoss               ultralytics/models/utils/loss.py:14-373
  RTDETRDetectionLoss    ultralytics/models/utils/loss.py:376-443
  VarifocalLoss / FocalLoss  ultralytics/utils/loss.py:135-178
  bbox_iou (IoU and RIOU branches)  ultralytics/utils/metrics.py:71-130

The reference builds the cost matrix on the GPU, copies it to the host and runs scipy.optimize.linear_sum_assignment
per image (ops.py:116-117) -- once per decoder layer plus once for the encoder proposals, i.e. four host round trips
in the middle of every training step.  Here the assignment is `tamtr_linear_sum_assignment` (csrc/assign.cu: the
same shortest-augmenting-path algorithm, fp64 arithmetic and tie order as SciPy, one warp per image), all layers in
one launch, and every index tensor that depends only on the batch's ground-truth counts is built once per distinct
`gt_groups` -- no synchronisation, CUDA-graph capturable.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

__all__ = ("HungarianMatcher", "DETRLoss", "RTDETRDetectionLoss", "VarifocalLoss", "FocalLoss", "bbox_iou",
           "linear_sum_assignment")


def bbox_iou(box1, box2, xywh=True, RIOU=False, eps=1e-7):
    """IoU / RIoU of broadcastable boxes (metrics.py:71-130, the two branches this path uses), same op order."""
    if xywh:
        (x1, y1, w1, h1), (x2, y2, w2, h2) = box1.chunk(4, -1), box2.chunk(4, -1)
        w1_, h1_, w2_, h2_ = w1 / 2, h1 / 2, w2 / 2, h2 / 2
        b1_x1, b1_x2, b1_y1, b1_y2 = x1 - w1_, x1 + w1_, y1 - h1_, y1 + h1_
        b2_x1, b2_x2, b2_y1, b2_y2 = x2 - w2_, x2 + w2_, y2 - h2_, y2 + h2_
    else:
        b1_x1, b1_y1, b1_x2, b1_y2 = box1.chunk(4, -1)
        b2_x1, b2_y1, b2_x2, b2_y2 = box2.chunk(4, -1)
        w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
        w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (b1_x2.minimum(b2_x2) - b1_x1.maximum(b2_x1)).clamp_(0) * \
            (b1_y2.minimum(b2_y2) - b1_y1.maximum(b2_y1)).clamp_(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    if not RIOU:
        return iou
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
    maxwh1 = torch.max(w1, h1)
    maxwh2 = torch.max(w2, h2)
    c2 = (maxwh1 + maxwh2 + torch.sqrt(rho2) + eps).pow(2)
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


class VarifocalLoss(nn.Module):
    @staticmethod
    def forward(pred_score, gt_score, label, alpha=0.75, gamma=2.0):
        weight = alpha * pred_score.sigmoid().pow(gamma) * (1 - label) + gt_score * label
        with torch.autocast("cuda", enabled=False):
            loss = (F.binary_cross_entropy_with_logits(pred_score.float(), gt_score.float(), reduction='none') *
                    weight).mean(1).sum()
        return loss


class FocalLoss(nn.Module):
    @staticmethod
    def forward(pred, label, gamma=1.5, alpha=0.25):
        loss = F.binary_cross_entropy_with_logits(pred, label, reduction='none')
        pred_prob = pred.sigmoid()
        p_t = label * pred_prob + (1 - label) * (1 - pred_prob)
        loss = loss * (1.0 - p_t) ** gamma
        if alpha > 0:
            loss = loss * (label * alpha + (1 - label) * (1 - alpha))
        return loss.mean(1).sum()


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


class HungarianMatcher(nn.Module):
    """ops.py:12-121.  forward() keeps the reference's signature and return value (a list of (query idx, gt idx) per
    image); match_layers() is the batched form the loss uses (all decoder layers in one launch)."""

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

    def cost_matrix(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls):
        """[..., nq, 4], [..., nq, nc] -> [..., nq, total_gt] (ops.py:77-112, same op order)."""
        pred_scores = pred_scores.detach()
        pred_scores = F.sigmoid(pred_scores) if self.use_fl else F.softmax(pred_scores, dim=-1)
        pred_bboxes = pred_bboxes.detach()
        pred_scores = pred_scores[..., gt_cls]
        if self.use_fl:
            neg_cost_class = (1 - self.alpha) * (pred_scores ** self.gamma) * (-(1 - pred_scores + 1e-8).log())
            pos_cost_class = self.alpha * ((1 - pred_scores) ** self.gamma) * (-(pred_scores + 1e-8).log())
            cost_class = pos_cost_class - neg_cost_class
        else:
            cost_class = -pred_scores
        cost_bbox = (pred_bboxes.unsqueeze(-2) - gt_bboxes).abs().sum(-1)
        cost_giou = 1.0 - bbox_iou(pred_bboxes.unsqueeze(-2), gt_bboxes, xywh=True, RIOU=True).squeeze(-1)
        C = self.cost_gain['class'] * cost_class + self.cost_gain['bbox'] * cost_bbox + self.cost_gain['giou'] * cost_giou
        return torch.where(torch.isfinite(C), C, torch.zeros((), dtype=C.dtype, device=C.device))

    def cost_matrix_per_image(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups):
        """[n_layers, bs, nq, .] -> [n_layers, bs, nq, max_gt]: only each image's OWN ground truths (the reference's
        [bs*nq, total_gt] matrix prices every query against every image's boxes and then throws all but one block per
        image away, ops.py:104-116).  Element-wise the same arithmetic, so the kept entries are bit-identical."""
        grp = _Groups.get(gt_groups, pred_scores.shape[2], pred_scores.device)
        gtb = gt_bboxes[grp.pad_index].unsqueeze(1)                                    # [bs, 1, max_gt, 4]
        cls = gt_cls[grp.pad_index]                                                    # [bs, max_gt]
        p = pred_scores.detach()
        p = F.sigmoid(p) if self.use_fl else F.softmax(p, dim=-1)
        n_l, bs, nq = p.shape[:3]
        p = torch.gather(p, 3, cls.view(1, bs, 1, -1).expand(n_l, bs, nq, -1))
        box = pred_bboxes.detach()
        if self.use_fl:
            neg_cost_class = (1 - self.alpha) * (p ** self.gamma) * (-(1 - p + 1e-8).log())
            pos_cost_class = self.alpha * ((1 - p) ** self.gamma) * (-(p + 1e-8).log())
            cost_class = pos_cost_class - neg_cost_class
        else:
            cost_class = -p
        cost_bbox = (box.unsqueeze(-2) - gtb).abs().sum(-1)
        cost_giou = 1.0 - bbox_iou(box.unsqueeze(-2), gtb, xywh=True, RIOU=True).squeeze(-1)
        C = self.cost_gain['class'] * cost_class + self.cost_gain['bbox'] * cost_bbox + self.cost_gain['giou'] * cost_giou
        return torch.where(torch.isfinite(C), C, torch.zeros((), dtype=C.dtype, device=C.device))

    def match_layers(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups):
        """pred_* [n_layers, bs, nq, .] -> (image idx [P], query idx [n_layers, P], gt idx [n_layers, P])."""
        if sum(gt_groups) == 0:
            z = torch.zeros(0, dtype=torch.long, device=pred_bboxes.device)
            e = torch.zeros(pred_bboxes.shape[0], 0, dtype=torch.long, device=pred_bboxes.device)
            return z, e, e.clone()
        C = self.cost_matrix_per_image(pred_bboxes.float(), pred_scores.float(), gt_bboxes.float(), gt_cls, gt_groups)
        return linear_sum_assignment(C, gt_groups, padded=True)

    def forward(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, masks=None, gt_mask=None):
        bs, nq, nc = pred_scores.shape
        if sum(gt_groups) == 0:
            return [(torch.tensor([], dtype=torch.long), torch.tensor([], dtype=torch.long)) for _ in range(bs)]
        _, q, g = self.match_layers(pred_bboxes.unsqueeze(0), pred_scores.unsqueeze(0), gt_bboxes, gt_cls, gt_groups)
        counts = _Groups.get(gt_groups, nq, pred_scores.device).counts
        return list(zip(q[0].split(counts), g[0].split(counts)))


class DETRLoss(nn.Module):
    """loss.py:14-373 (focal / varifocal classification loss, L1 + RIoU box losses, auxiliary losses per layer)."""

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
        self.fl = FocalLoss() if use_fl else None
        self.vfl = VarifocalLoss() if use_vfl else None
        self.use_uni_match = use_uni_match
        self.uni_match_ind = uni_match_ind
        self.device = None

    # ---- per-layer pieces (loss.py:85-167, 282-326) on pre-matched pairs
    def _get_loss_class(self, pred_scores, targets, gt_scores, num_gts, postfix=''):
        bs, nq = pred_scores.shape[:2]
        one_hot = torch.zeros((bs, nq, self.nc + 1), dtype=torch.int64, device=targets.device)
        one_hot.scatter_(2, targets.unsqueeze(-1), 1)
        one_hot = one_hot[..., :-1]
        gt_scores = gt_scores.view(bs, nq, 1) * one_hot
        if self.fl:
            if num_gts and self.vfl:
                loss_cls = self.vfl(pred_scores, gt_scores, one_hot)
            else:
                loss_cls = self.fl(pred_scores, one_hot.float())
            loss_cls = loss_cls / (max(num_gts, 1) / nq)
        else:
            loss_cls = nn.BCEWithLogitsLoss(reduction='none')(pred_scores, gt_scores).mean(1).sum()
        return {f'loss_class{postfix}': loss_cls.squeeze() * self.loss_gain['class']}

    def _get_loss_bbox(self, pred_bboxes, gt_bboxes, postfix=''):
        name_bbox, name_giou = f'loss_bbox{postfix}', f'loss_giou{postfix}'
        if len(gt_bboxes) == 0:
            z = torch.zeros((), device=self.device)       # (a fill kernel, not a host->device copy: graph-capturable)
            return {name_bbox: z, name_giou: z.clone()}
        loss = {name_bbox: self.loss_gain['bbox'] * F.l1_loss(pred_bboxes, gt_bboxes, reduction='sum') / len(gt_bboxes)}
        giou = 1.0 - bbox_iou(pred_bboxes, gt_bboxes, xywh=True, RIOU=True)
        loss[name_giou] = self.loss_gain['giou'] * (giou.sum() / len(gt_bboxes))
        return {k: v.squeeze() for k, v in loss.items()}

    def _get_loss(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, masks=None, gt_mask=None, postfix='',
                  match_indices=None):
        """One layer.  match_indices: None (match here), the reference's list of (query idx, gt idx) per image, or the
        flat triple (image idx, query idx, gt idx)."""
        if match_indices is None:
            img, q, g = self.matcher.match_layers(pred_bboxes.unsqueeze(0), pred_scores.unsqueeze(0), gt_bboxes, gt_cls,
                                                  gt_groups)
            match_indices = (img, q[0], g[0])
        if isinstance(match_indices, list):
            dev = pred_bboxes.device
            img = torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(match_indices)]).to(dev)
            match_indices = (img, torch.cat([s for s, _ in match_indices]).to(dev),
                             torch.cat([d for _, d in match_indices]).to(dev))
        img, src, gt_idx = match_indices
        idx = (img, src)
        bs, nq = pred_bboxes.shape[:2]
        pred_bboxes, gt_bboxes = pred_bboxes[idx], gt_bboxes[gt_idx]
        if pred_scores is None:
            return dict(self._get_loss_bbox(pred_bboxes, gt_bboxes, postfix))
        targets = torch.full((bs, nq), self.nc, device=pred_scores.device, dtype=gt_cls.dtype)
        targets[idx] = gt_cls[gt_idx]
        gt_scores = torch.zeros([bs, nq], device=pred_scores.device)
        if len(gt_bboxes):
            gt_scores[idx] = bbox_iou(pred_bboxes.detach(), gt_bboxes, xywh=True).squeeze(-1)
        loss = {}
        loss.update(self._get_loss_class(pred_scores, targets, gt_scores, len(gt_bboxes), postfix))
        loss.update(self._get_loss_bbox(pred_bboxes, gt_bboxes, postfix))
        return loss

    def _get_loss_aux(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, match_indices=None, postfix='',
                      masks=None, gt_mask=None):
        loss = torch.zeros(3, device=pred_bboxes.device)
        if match_indices is None and self.use_uni_match:
            match_indices = self.matcher(pred_bboxes[self.uni_match_ind], pred_scores[self.uni_match_ind], gt_bboxes,
                                         gt_cls, gt_groups)
        for i, aux_bboxes in enumerate(pred_bboxes):
            aux_scores = None if pred_scores is None else pred_scores[i]
            loss_ = self._get_loss(aux_bboxes, aux_scores, gt_bboxes, gt_cls, gt_groups, postfix=postfix,
                                   match_indices=match_indices)
            if aux_scores is not None:
                loss[0] = loss[0] + loss_[f'loss_class{postfix}']
            loss[1] = loss[1] + loss_[f'loss_bbox{postfix}']
            loss[2] = loss[2] + loss_[f'loss_giou{postfix}']
        return {f'loss_class_aux{postfix}': loss[0], f'loss_bbox_aux{postfix}': loss[1], f'loss_giou_aux{postfix}': loss[2]}

    def _get_loss_layers(self, pred_bboxes, pred_scores, gt_bboxes, gt_cls, img, q, g, postfix=''):
        """All layers at once (the reference loops, loss.py:232-243; same arithmetic per layer, one set of kernels):
        pred_* [n_l, bs, nq, .], img [P], q / g [n_l, P] (or [P] for the fixed denoising matches).  Returns the loss dict
        of loss.py:328-373: last layer + the sum over the auxiliary layers."""
        n_l, bs, nq = pred_bboxes.shape[:3]
        P = img.numel()
        if q.dim() == 1:
            q, g = q.expand(n_l, P), g.expand(n_l, P)
        lay = torch.arange(n_l, device=img.device).unsqueeze(1).expand(n_l, P)
        idx = (lay, img.expand(n_l, P), q)
        pb, gb = pred_bboxes[idx], gt_bboxes[g]                                        # [n_l, P, 4]
        bbox = self.loss_gain['bbox'] * (pb - gb).abs().sum((1, 2)) / P
        giou = self.loss_gain['giou'] * ((1.0 - bbox_iou(pb, gb, xywh=True, RIOU=True)).sum((1, 2)) / P)
        out = {}
        if pred_scores is not None:
            targets = torch.full((n_l, bs, nq), self.nc, device=pred_scores.device, dtype=gt_cls.dtype)
            targets[idx] = gt_cls[g]
            gt_scores = torch.zeros((n_l, bs, nq), device=pred_scores.device)
            gt_scores[idx] = bbox_iou(pb.detach(), gb, xywh=True).squeeze(-1)
            one_hot = torch.zeros((n_l, bs, nq, self.nc + 1), dtype=torch.int64, device=targets.device)
            one_hot.scatter_(3, targets.unsqueeze(-1), 1)
            one_hot = one_hot[..., :-1]
            gt_s = gt_scores.unsqueeze(-1) * one_hot
            if self.fl:
                if self.vfl:
                    weight = 0.75 * pred_scores.sigmoid().pow(2.0) * (1 - one_hot) + gt_s * one_hot
                    with torch.autocast("cuda", enabled=False):
                        cls = (F.binary_cross_entropy_with_logits(pred_scores.float(), gt_s.float(), reduction='none')
                               * weight).mean(2).sum((1, 2))
                else:
                    label = one_hot.float()
                    l = F.binary_cross_entropy_with_logits(pred_scores, label, reduction='none')
                    pr = pred_scores.sigmoid()
                    p_t = label * pr + (1 - label) * (1 - pr)
                    cls = (l * (1.0 - p_t) ** 1.5 * (label * 0.25 + (1 - label) * 0.75)).mean(2).sum((1, 2))
                cls = cls / (max(P, 1) / nq)
            else:
                cls = nn.BCEWithLogitsLoss(reduction='none')(pred_scores, gt_s).mean(2).sum((1, 2))
            cls = cls * self.loss_gain['class']
            out[f'loss_class{postfix}'] = cls[-1]
        out[f'loss_bbox{postfix}'] = bbox[-1]
        out[f'loss_giou{postfix}'] = giou[-1]
        if self.aux_loss:
            zero = torch.zeros((), device=pred_bboxes.device)
            out[f'loss_class_aux{postfix}'] = cls[:-1].sum() if pred_scores is not None else zero
            out[f'loss_bbox_aux{postfix}'] = bbox[:-1].sum()
            out[f'loss_giou_aux{postfix}'] = giou[:-1].sum()
        return out

    def forward(self, pred_bboxes, pred_scores, batch, postfix='', **kwargs):
        """pred_bboxes [l, b, query, 4], pred_scores [l, b, query, nc] (or None); batch: cls / bboxes / gt_groups."""
        _lib.require_cuda(pred_bboxes)
        self.device = pred_bboxes.device
        match_indices = kwargs.get('match_indices', None)
        gt_cls, gt_bboxes, gt_groups = batch['cls'], batch['bboxes'], batch['gt_groups']
        batched = pred_bboxes.is_cuda and not self.use_uni_match and sum(gt_groups) > 0
        if batched and isinstance(match_indices, tuple) and match_indices[0].numel() > 0:
            # fixed matches (denoising queries): the same pairs for every layer
            pb = pred_bboxes if self.aux_loss else pred_bboxes[-1:]
            ps = pred_scores if (self.aux_loss or pred_scores is None) else pred_scores[-1:]
            return self._get_loss_layers(pb, ps, gt_bboxes, gt_cls, *match_indices, postfix=postfix)
        if batched and match_indices is None and pred_scores is not None:
            # every layer's assignment in ONE launch (the reference matches layer by layer, loss.py:293-300, 232-243)
            n_l = pred_bboxes.shape[0] if self.aux_loss else 1
            img, q, g = self.matcher.match_layers(pred_bboxes[-n_l:], pred_scores[-n_l:], gt_bboxes, gt_cls, gt_groups)
            if img.numel() > 0:
                return self._get_loss_layers(pred_bboxes[-n_l:], pred_scores[-n_l:], gt_bboxes, gt_cls, img, q, g,
                                             postfix=postfix)
        total_loss = self._get_loss(pred_bboxes[-1], None if pred_scores is None else pred_scores[-1], gt_bboxes, gt_cls,
                                    gt_groups, postfix=postfix, match_indices=match_indices)
        if self.aux_loss:
            total_loss.update(self._get_loss_aux(pred_bboxes[:-1], None if pred_scores is None else pred_scores[:-1],
                                                 gt_bboxes, gt_cls, gt_groups, match_indices, postfix))
        return total_loss


class RTDETRDetectionLoss(DETRLoss):
    """loss.py:376-443: the detection loss plus the denoising loss on the CDN queries (fixed matches)."""

    def forward(self, preds, batch, dn_bboxes=None, dn_scores=None, dn_meta=None):
        # (explicit DETRLoss.forward instead of super(): patch.enable() binds this function onto the reference's class)
        pred_bboxes, pred_scores = preds
        total_loss = DETRLoss.forward(self, pred_bboxes, pred_scores, batch)
        if dn_meta is not None:
            dn_pos_idx, dn_num_group = dn_meta['dn_pos_idx'], dn_meta['dn_num_group']
            assert len(batch['gt_groups']) == len(dn_pos_idx)
            # the denoising matches depend only on the batch: flat device copy built once per dn_meta (no per-step
            # host->device index copies, which would also break CUDA-graph capture)
            cached = dn_meta.get('_tamtr_dn_match')
            if cached is None or cached[0].device != dn_bboxes.device:
                mi = RTDETRDetectionLoss.get_dn_match_indices(dn_pos_idx, dn_num_group, batch['gt_groups'])
                dev = dn_bboxes.device
                cached = (torch.cat([torch.full_like(src, i) for i, (src, _) in enumerate(mi)]).long().to(dev),
                          torch.cat([src for src, _ in mi]).long().to(dev), torch.cat([dst for _, dst in mi]).to(dev))
                dn_meta['_tamtr_dn_match'] = cached
            dn_loss = DETRLoss.forward(self, dn_bboxes, dn_scores, batch, postfix='_dn', match_indices=cached)
            total_loss.update(dn_loss)
        else:
            total_loss.update({f'{k}_dn': torch.zeros((), device=self.device) for k in total_loss.keys()})
        return total_loss

    @staticmethod
    def get_dn_match_indices(dn_pos_idx, dn_num_group, gt_groups):
        dn_match_indices = []
        idx_groups = torch.as_tensor([0, *gt_groups[:-1]]).cumsum_(0)
        for i, num_gt in enumerate(gt_groups):
            if num_gt > 0:
                gt_idx = torch.arange(end=num_gt, dtype=torch.long) + idx_groups[i]
                gt_idx = gt_idx.repeat(dn_num_group)
                assert len(dn_pos_idx[i]) == len(gt_idx), 'Expected the same length'
                dn_match_indices.append((dn_pos_idx[i], gt_idx.to(dn_pos_idx[i].device)))
            else:
                dn_match_indices.append((torch.zeros([0], dtype=torch.long), torch.zeros([0], dtype=torch.long)))
        return dn_match_indices
