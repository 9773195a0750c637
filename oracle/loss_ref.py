"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's detection loss and Hungarian matching.

Follows (file:line under /root/reference/ultralytics):
  HungarianMatcher.forward     models/utils/ops.py:48-121  (cost matrix, scipy.optimize.linear_sum_assignment per image)
  DETRLoss                     models/utils/loss.py:85-167 (class / box losses), :282-326 (_get_loss), :199-263 (aux)
  RTDETRDetectionLoss.forward  models/utils/loss.py:384-443 (denoising part with fixed matches)
  VarifocalLoss / FocalLoss    utils/loss.py:135-178
  bbox_iou (IoU, RIOU)         utils/metrics.py:71-130
as configured by TAM-TR (nn/tasks.py:578: use_vfl=True, focal cost in the matcher, cost gains class 2 / bbox 5 /
giou 2, loss gains class 1 / bbox 5 / giou 2).  Written flat (functions over tensors), independent of the product.
Pinned to the reference's own outputs by tests/golden/loss.pt (oracle/make_goldens_loss.py).
"""
import math

import torch
import torch.nn.functional as F
from scipy.optimize import linear_sum_assignment

NC_DEFAULT = 10
COST_GAIN = {"class": 2.0, "bbox": 5.0, "giou": 2.0}
LOSS_GAIN = {"class": 1.0, "bbox": 5.0, "giou": 2.0}


def iou_xywh(b1, b2, riou=False, eps=1e-7):
    (x1, y1, w1, h1), (x2, y2, w2, h2) = b1.chunk(4, -1), b2.chunk(4, -1)
    ax1, ax2, ay1, ay2 = x1 - w1 / 2, x1 + w1 / 2, y1 - h1 / 2, y1 + h1 / 2
    bx1, bx2, by1, by2 = x2 - w2 / 2, x2 + w2 / 2, y2 - h2 / 2, y2 + h2 / 2
    inter = (ax2.minimum(bx2) - ax1.maximum(bx1)).clamp(0) * (ay2.minimum(by2) - ay1.maximum(by1)).clamp(0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    if not riou:
        return iou
    rho2 = ((bx1 + bx2 - ax1 - ax2) ** 2 + (by1 + by2 - ay1 - ay2) ** 2) / 4
    c2 = (torch.max(w1, h1) + torch.max(w2, h2) + torch.sqrt(rho2) + eps).pow(2)
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    alpha = (v / (v - iou + (1 + eps))).detach()
    return iou - (rho2 / c2 + v * alpha)


def cost_matrix(pred_bboxes, pred_scores, gt_bboxes, gt_cls, alpha=0.25, gamma=2.0):
    """[bs, nq, 4], [bs, nq, nc] -> [bs, nq, total_gt]"""
    p = pred_scores.detach().sigmoid()[..., gt_cls]
    neg = (1 - alpha) * (p ** gamma) * (-(1 - p + 1e-8).log())
    pos = alpha * ((1 - p) ** gamma) * (-(p + 1e-8).log())
    box = pred_bboxes.detach()
    cost_bbox = (box.unsqueeze(-2) - gt_bboxes).abs().sum(-1)
    cost_giou = 1.0 - iou_xywh(box.unsqueeze(-2), gt_bboxes, riou=True).squeeze(-1)
    C = COST_GAIN["class"] * (pos - neg) + COST_GAIN["bbox"] * cost_bbox + COST_GAIN["giou"] * cost_giou
    C[C.isnan() | C.isinf()] = 0.0
    return C


def hungarian_match(pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups):
    """-> (image idx, query idx, global gt idx), pairs of each image in ascending query order."""
    bs = pred_bboxes.shape[0]
    if sum(gt_groups) == 0:
        z = torch.zeros(0, dtype=torch.long)
        return z, z.clone(), z.clone()
    C = cost_matrix(pred_bboxes, pred_scores, gt_bboxes, gt_cls).cpu()
    img, qs, gs, start = [], [], [], 0
    for b, (c, n) in enumerate(zip(C.split(list(gt_groups), -1), gt_groups)):
        i, j = linear_sum_assignment(c[b].numpy())
        img.append(torch.full((len(i),), b, dtype=torch.long))
        qs.append(torch.as_tensor(i, dtype=torch.long))
        gs.append(torch.as_tensor(j, dtype=torch.long) + start)
        start += n
    return torch.cat(img), torch.cat(qs), torch.cat(gs)


def layer_loss(pred_bboxes, pred_scores, gt_bboxes, gt_cls, match, nc, postfix=""):
    img, q, g = match
    bs, nq = pred_bboxes.shape[:2]
    pb, gb = pred_bboxes[img, q], gt_bboxes[g]
    out = {}
    n = len(gb)
    if pred_scores is not None:
        targets = torch.full((bs, nq), nc, dtype=gt_cls.dtype)
        targets[img, q] = gt_cls[g]
        gt_scores = torch.zeros(bs, nq)
        if n:
            gt_scores[img, q] = iou_xywh(pb.detach(), gb).squeeze(-1)
        one_hot = F.one_hot(targets, nc + 1)[..., :-1]
        gt_s = gt_scores.view(bs, nq, 1) * one_hot
        if n:   # varifocal
            weight = 0.75 * pred_scores.sigmoid().pow(2.0) * (1 - one_hot) + gt_s * one_hot
            cls = (F.binary_cross_entropy_with_logits(pred_scores.float(), gt_s.float(), reduction="none") * weight).mean(1).sum()
        else:   # focal
            label = one_hot.float()
            l = F.binary_cross_entropy_with_logits(pred_scores, label, reduction="none")
            pr = pred_scores.sigmoid()
            p_t = label * pr + (1 - label) * (1 - pr)
            cls = (l * (1.0 - p_t) ** 1.5 * (label * 0.25 + (1 - label) * 0.75)).mean(1).sum()
        out[f"loss_class{postfix}"] = cls / (max(n, 1) / nq) * LOSS_GAIN["class"]
    if n == 0:
        out[f"loss_bbox{postfix}"] = torch.tensor(0.0)
        out[f"loss_giou{postfix}"] = torch.tensor(0.0)
    else:
        out[f"loss_bbox{postfix}"] = LOSS_GAIN["bbox"] * (pb - gb).abs().sum() / n
        out[f"loss_giou{postfix}"] = LOSS_GAIN["giou"] * ((1.0 - iou_xywh(pb, gb, riou=True)).sum() / n)
    return out


def detr_loss(pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, nc, postfix="", match=None):
    """pred_* [l, bs, nq, .]: last layer + auxiliary layers; match: fixed (img, q, g) or None (Hungarian per layer)."""
    def m(l):
        return match if match is not None else hungarian_match(pred_bboxes[l], pred_scores[l], gt_bboxes, gt_cls, gt_groups)
    total = layer_loss(pred_bboxes[-1], pred_scores[-1], gt_bboxes, gt_cls, m(-1), nc, postfix)
    aux = [torch.zeros(()), torch.zeros(()), torch.zeros(())]
    for l in range(pred_bboxes.shape[0] - 1):
        d = layer_loss(pred_bboxes[l], pred_scores[l], gt_bboxes, gt_cls, m(l), nc, postfix)
        aux = [aux[0] + d[f"loss_class{postfix}"], aux[1] + d[f"loss_bbox{postfix}"], aux[2] + d[f"loss_giou{postfix}"]]
    total.update({f"loss_class_aux{postfix}": aux[0], f"loss_bbox_aux{postfix}": aux[1], f"loss_giou_aux{postfix}": aux[2]})
    return total


def dn_match(dn_pos_idx, dn_num_group, gt_groups):
    img, q, g, start = [], [], [], 0
    for b, n in enumerate(gt_groups):
        if n > 0:
            idx = dn_pos_idx[b].long().cpu()
            img.append(torch.full((len(idx),), b, dtype=torch.long))
            q.append(idx)
            g.append((torch.arange(n) + start).repeat(dn_num_group))
        start += n
    z = torch.zeros(0, dtype=torch.long)
    return (torch.cat(img), torch.cat(q), torch.cat(g)) if img else (z, z.clone(), z.clone())


def rtdetr_detection_loss(pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, nc=NC_DEFAULT, dn_bboxes=None,
                          dn_scores=None, dn_meta=None):
    total = detr_loss(pred_bboxes, pred_scores, gt_bboxes, gt_cls, gt_groups, nc)
    if dn_meta is not None:
        match = dn_match(dn_meta["dn_pos_idx"], dn_meta["dn_num_group"], gt_groups)
        total.update(detr_loss(dn_bboxes, dn_scores, gt_bboxes, gt_cls, gt_groups, nc, "_dn", match))
    else:
        total.update({f"{k}_dn": torch.tensor(0.0) for k in list(total)})
    return total


def make_case(seed, n_layers, bs, nq, nc, gt_groups, dn_groups=0, tie_boxes=False):
    """Seeded inputs of a loss case (shared by the golden generator and the tests)."""
    from oracle import seeding
    G = sum(gt_groups)
    boxes = torch.cat([seeding.seeded_uniform(seed, "gt_xy", (G, 2), 0.1, 0.9),
                       seeding.seeded_uniform(seed, "gt_wh", (G, 2), 0.02, 0.3)], -1)
    cls = (seeding.seeded_uniform(seed, "gt_cls", (G,)) * nc).long().clamp(max=nc - 1)
    pb = torch.cat([seeding.seeded_uniform(seed, "p_xy", (n_layers, bs, nq, 2), 0.05, 0.95),
                    seeding.seeded_uniform(seed, "p_wh", (n_layers, bs, nq, 2), 0.02, 0.4)], -1)
    if tie_boxes:      # identical predictions everywhere: every cost column is constant -> SciPy's tie order decides
        pb = pb[:, :, :1].expand(-1, -1, nq, -1).contiguous()
    ps = seeding.seeded_tensor(seed, "p_s", (n_layers, bs, nq, nc)) * (0.0 if tie_boxes else 2.0)
    case = {"pred_bboxes": pb, "pred_scores": ps, "gt_bboxes": boxes, "gt_cls": cls, "gt_groups": list(gt_groups), "nc": nc}
    if dn_groups:
        max_gt = max(gt_groups)
        n_dn = 2 * max_gt * dn_groups
        pos = []
        for n in gt_groups:     # positive copies first in each group, like ops.py:224-233
            idx = torch.cat([torch.arange(n) + 2 * max_gt * k for k in range(dn_groups)]) if n else torch.zeros(0, dtype=torch.long)
            pos.append(idx.long())
        case["dn_meta"] = {"dn_pos_idx": pos, "dn_num_group": dn_groups, "dn_num_split": [n_dn, nq]}
        case["dn_bboxes"] = torch.cat([seeding.seeded_uniform(seed, "d_xy", (n_layers - 1, bs, n_dn, 2), 0.05, 0.95),
                                       seeding.seeded_uniform(seed, "d_wh", (n_layers - 1, bs, n_dn, 2), 0.02, 0.4)], -1)
        case["dn_scores"] = seeding.seeded_tensor(seed, "d_s", (n_layers - 1, bs, n_dn, nc)) * 2.0
    return case


CASES = {
    "visdrone_like": dict(seed=301, n_layers=4, bs=4, nq=100, nc=10, gt_groups=[23, 0, 57, 100], dn_groups=1),
    "more_gt_than_queries": dict(seed=302, n_layers=2, bs=3, nq=20, nc=10, gt_groups=[35, 20, 7], dn_groups=0),
    "ties": dict(seed=303, n_layers=2, bs=2, nq=40, nc=10, gt_groups=[12, 30], dn_groups=0, tie_boxes=True),
    "no_targets": dict(seed=304, n_layers=2, bs=2, nq=30, nc=10, gt_groups=[0, 0], dn_groups=0),
}
