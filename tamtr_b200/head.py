"""Detection heads around the hot path: host-side glue mirroring the reference's heads.

  RTDETRDecoder      ultralytics/nn/modules/head.py:174-435
  ManbaWorldDecoder  ultralytics/nn/modules/head.py:1005-1290   (the head TAMTR.yaml:67 builds: "MEH")
  get_cdn_group      ultralytics/models/utils/ops.py:152-291    (contrastive denoising queries)

Same constructor arguments, forward signatures, outputs and state_dict keys.  The glue (1x1 input projection,
anchors, top-k query selection, denoising group) stays a sequence of library ops as in the reference; the decoder it
drives runs on the sm_100a kernels (modules.py).  ManbaWorldDecoder's VSSBlocks (VMamba selective scan; the reference
needs an external CUDA extension for them that is not part of its tree -- SURVEY.md section 8c) run on the kernels of
csrc/sscan.cu (vss.py); `vss=False` replaces them by identities, the configuration of the head-level parity fixtures.
"""
import math

import os

import torch
import torch.nn as nn

from . import fold, ops
from .loss import DeviceTargets
from .vss import VSSBlock
from .modules import (MLP, ContrastiveHeadMLP, DeformableTransformerDecoder, DeformableTransformerDecoderLayer,
                      TextDeformableTransformerDecoder, is_plain_msda)

__all__ = ("RTDETRDecoder", "ManbaWorldDecoder", "get_cdn_group", "plan_cdn_group", "CdnPlan", "cdn_group_device")


def _xywh_to_xyxy(b):
    half = b[..., 2:] / 2
    return torch.cat([b[..., :2] - half, b[..., :2] + half], -1)


def _xyxy_to_xywh(b):
    return torch.cat([(b[..., :2] + b[..., 2:]) / 2, b[..., 2:] - b[..., :2]], -1)


class CdnPlan:
    """Host-side result of the denoising-group construction: everything except the (learnable) class-embedding
    lookup.  Splitting it off lets a training step be captured in a CUDA graph: the plan is built per batch on the
    host exactly as the reference does, the graph only contains the embedding gather."""

    def __init__(self, dn_cls, dn_box, dn_img, slot, mask, meta, bs, n_dn):
        self.dn_cls, self.dn_box, self.dn_img, self.slot = dn_cls, dn_box, dn_img, slot
        self.mask, self.meta, self.bs, self.n_dn = mask, meta, bs, n_dn

    def to(self, device):
        return CdnPlan(self.dn_cls.to(device), self.dn_box.to(device), self.dn_img.to(device).long(),
                       self.slot.to(device), self.mask.to(device), self.meta, self.bs, self.n_dn)

    def materialize(self, class_embed):
        dev = class_embed.device
        embed = ops.embed_rows(class_embed, self.dn_cls.to(dev))
        # (the reference fills fp32 buffers on the targets' device and moves them afterwards, ops.py:243-262; filling
        #  them on the embedding's device is the same result without a host round trip)
        pad_embed = torch.zeros(self.bs, self.n_dn, embed.shape[-1], device=dev, dtype=embed.dtype)
        pad_box = torch.zeros(self.bs, self.n_dn, 4, device=dev, dtype=self.dn_box.dtype)
        where = (self.dn_img.to(dev).long(), self.slot.to(dev))
        pad_embed[where] = embed
        pad_box[where] = self.dn_box.to(dev)
        return pad_embed, pad_box, self.mask.to(dev), self.meta


def plan_cdn_group(batch, num_classes, num_queries, num_dn=100, cls_noise_ratio=0.5, box_noise_scale=1.0,
                   training=False):
    """Host part of get_cdn_group (ops.py:152-291).  Consumes the global RNG in the same order as the reference
    (label mask, replacement labels, box-noise sign, box-noise magnitude), so seeded runs produce the same queries.
    Returns a CdnPlan or None."""
    if (not training) or num_dn <= 0:
        return None
    groups = batch["gt_groups"]
    total, biggest = sum(groups), max(groups)
    if biggest == 0:
        return None
    n_group = max(1, num_dn // biggest)
    bs = len(groups)
    gt_cls, gt_box, gt_img = batch["cls"], batch["bboxes"], batch["batch_idx"]

    # every group holds one positive and one negative copy of each ground truth
    dn_cls = gt_cls.repeat(2 * n_group)
    dn_box = gt_box.repeat(2 * n_group, 1)
    dn_img = gt_img.repeat(2 * n_group).view(-1)
    negatives = torch.arange(total * n_group, dtype=torch.long, device=gt_box.device) + n_group * total

    if cls_noise_ratio > 0:
        flip = torch.rand(dn_cls.shape) < (cls_noise_ratio * 0.5)
        where = torch.nonzero(flip).squeeze(-1)
        dn_cls[where] = torch.randint_like(where, 0, num_classes, dtype=dn_cls.dtype, device=dn_cls.device)

    if box_noise_scale > 0:
        corners = _xywh_to_xyxy(dn_box)
        span = (dn_box[..., 2:] * 0.5).repeat(1, 2) * box_noise_scale
        sign = torch.randint_like(dn_box, 0, 2) * 2.0 - 1.0
        mag = torch.rand_like(dn_box)
        mag[negatives] += 1.0            # negatives are pushed 1x..2x the half-size away
        mag *= sign
        corners += mag * span
        corners.clip_(min=0.0, max=1.0)
        dn_box = torch.logit(_xyxy_to_xywh(corners), eps=1e-6)

    n_dn = int(biggest * 2 * n_group)
    slot = torch.cat([torch.arange(n, dtype=torch.long) for n in groups])
    pos_idx = torch.stack([slot + biggest * i for i in range(n_group)], dim=0)
    slot = torch.cat([slot + biggest * i for i in range(2 * n_group)])

    size = n_dn + num_queries
    mask = torch.zeros([size, size], dtype=torch.bool)
    mask[n_dn:, :n_dn] = True                       # matching queries never see the denoising queries
    for i in range(n_group):                        # denoising groups never see each other
        lo, hi = biggest * 2 * i, biggest * 2 * (i + 1)
        mask[lo:hi, hi:n_dn] = True
        mask[lo:hi, :lo] = True
    meta = {"dn_pos_idx": [p.reshape(-1) for p in pos_idx.cpu().split(list(groups), dim=1)],
            "dn_num_group": n_group, "dn_num_split": [n_dn, num_queries]}
    return CdnPlan(dn_cls.long(), dn_box, dn_img, slot, mask, meta, bs, n_dn)


def get_cdn_group(batch, num_classes, num_queries, class_embed, num_dn=100, cls_noise_ratio=0.5,
                  box_noise_scale=1.0, training=False):
    """Contrastive denoising group, same signature and results as the reference (ops.py:152-291).

    batch: {'cls' [n], 'bboxes' [n,4] (cx,cy,w,h), 'batch_idx' [n], 'gt_groups' [B ints]}.
    Returns (dn_embed [B,num_dn,hd], dn_bbox [B,num_dn,4] (logit space), attn_mask [Lq,Lq] bool, dn_meta)."""
    plan = plan_cdn_group(batch, num_classes, num_queries, num_dn, cls_noise_ratio, box_noise_scale, training)
    if plan is None:
        return None, None, None, None
    return plan.materialize(class_embed)


def cdn_group_device(targets, num_classes, num_queries, class_embed, num_dn=100, cls_noise_ratio=0.5, box_noise_scale=1.0):
    """get_cdn_group (ops.py:152-291) for a batch held in fixed-shape device tensors (loss.DeviceTargets), as ONE kernel
    (tamtr_cdn_group) + the class-embedding gather: no host planning, no data-dependent shapes, so the training step can
    be captured once and replayed on every batch.  The group occupies the first 2 * max_gt * num_group of
    `targets.dn_capacity` slots exactly as the reference lays it out; the remaining slots of the bucket are padding
    (blocked in the attention mask, ignored by the loss), so the matching and denoising queries see what they see in the
    reference.

    Randomness: the reference draws from torch's global generator with data-dependent sizes (ops.py:217-229), which no
    fixed-shape program can reproduce draw for draw.  Contract here: ONE torch.rand([B, capacity, 10]) per step (graph-safe
    Philox); slot (b, s) consumes its ten numbers as (label-flip test, replacement label, 4 box-noise signs, 4 box-noise
    magnitudes) through the reference's formulas -- same distributions, and bit-identical queries for identical numbers
    (tests/test_cdn.py checks the formulas against the reference with its RNG calls replaced by the same numbers)."""
    from . import _lib
    _lib.require_cuda(targets.boxes, class_embed)
    B, G, D = targets.bs, targets.max_gt, targets.dn_capacity
    if D is None:
        raise RuntimeError("tamtr_b200: DeviceTargets.dn_capacity is not set (use DeviceTargets.from_batch / capacity_for)")
    dev = targets.boxes.device
    uni = torch.rand(B, D, 10, dtype=torch.float32, device=dev)
    dn_cls = torch.empty(B, D, dtype=torch.int64, device=dev)
    dn_box = torch.empty(B, D, 4, dtype=torch.float32, device=dev)
    valid = torch.empty(B, D, dtype=torch.float32, device=dev)
    Lq = D + num_queries
    mask = torch.empty(Lq, Lq, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().tamtr_cdn_group(targets.boxes.data_ptr(), targets.cls.data_ptr(), targets.count.data_ptr(),
                                        uni.data_ptr(), dn_cls.data_ptr(), dn_box.data_ptr(), valid.data_ptr(),
                                        mask.data_ptr(), B, G, D, num_queries, num_classes, int(num_dn),
                                        float(cls_noise_ratio), float(box_noise_scale), _lib.stream_ptr(dev))
    _lib.check(rc, "cdn_group")
    embed = ops.embed_rows(class_embed, dn_cls) * valid.unsqueeze(-1).to(class_embed.dtype)
    meta = {"dn_num_split": [D, num_queries], "num_dn_cfg": int(num_dn), "dn_valid": valid, "dn_cls": dn_cls}
    return embed, dn_box, mask.view(torch.bool), meta


class _HeadBase(nn.Module):
    """What RTDETRDecoder and ManbaWorldDecoder share (head.py:184-435 / 1015-1290)."""
    export = False

    def _build_common(self, nc, ch, hd, nq, nd, label_noise_ratio, box_noise_scale, learnt_init_query):
        self.hidden_dim = hd
        self.nl = len(ch)
        self.nc = nc
        self.num_queries = nq
        self.input_proj = nn.ModuleList(
            nn.Sequential(nn.Conv2d(c, hd, 1, bias=False), nn.BatchNorm2d(hd)) for c in ch)
        self.denoising_class_embed = nn.Embedding(nc + 1, hd)
        self.num_denoising = nd
        self.label_noise_ratio = label_noise_ratio
        self.box_noise_scale = box_noise_scale
        self.learnt_init_query = learnt_init_query
        if learnt_init_query:
            self.tgt_embed = nn.Embedding(nq, hd)
        self.query_pos_head = MLP(4, 2 * hd, hd, num_layers=2)
        self.enc_output = nn.Sequential(nn.Linear(hd, hd), nn.LayerNorm(hd))
        self.enc_score_head = nn.Linear(hd, nc)
        self.enc_bbox_head = MLP(hd, hd, 4, num_layers=3)

    @staticmethod
    def _generate_anchors(shapes, grid_size=0.05, dtype=torch.float32, device="cpu", eps=1e-2):
        """One anchor per pyramid cell: centre (x+0.5)/W, (y+0.5)/H, size grid_size * 2^level; returned in logit
        space with invalid (too close to the border) anchors set to +inf (head.py:1177-1200)."""
        per_level = []
        for lvl, (h, w) in enumerate(shapes):
            ys = torch.arange(end=h, dtype=dtype, device=device)
            xs = torch.arange(end=w, dtype=dtype, device=device)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            centre = (torch.stack([gx, gy], -1).unsqueeze(0) + 0.5) / torch.tensor([h, w], dtype=dtype, device=device)
            size = torch.ones_like(centre) * grid_size * (2.0 ** lvl)
            per_level.append(torch.cat([centre, size], -1).view(-1, h * w, 4))
        anchors = torch.cat(per_level, 1)
        valid = ((anchors > eps) * (anchors < 1 - eps)).all(-1, keepdim=True)
        anchors = torch.log(anchors / (1 - anchors)).masked_fill(~valid, float("inf"))
        return anchors, valid

    fused_input_proj = True

    def _fusable_input_proj(self, x):
        return (self.fused_input_proj and x[0].is_cuda and all(
            isinstance(p, nn.Sequential) and len(p) == 2 and isinstance(p[0], nn.Conv2d) and p[0].bias is None
            and p[0].kernel_size == (1, 1) and p[0].stride == (1, 1) and p[0].groups == 1
            and isinstance(p[1], nn.BatchNorm2d) and p[1].affine for p in self.input_proj))

    def _get_encoder_input(self, x):
        if self._fusable_input_proj(x):
            return ops.input_proj_tokens(list(x), self.input_proj, self.training)
        feats, shapes = [], []
        for proj, fmap in zip(self.input_proj, x):
            f = proj(fmap)
            shapes.append([f.shape[2], f.shape[3]])
            feats.append(f.flatten(2).permute(0, 2, 1))
        return torch.cat(feats, 1), shapes

    def _anchors(self, shapes, dtype, device):
        """Anchors depend only on the pyramid geometry: built once per (shapes, dtype, device) instead of on every
        forward (the reference rebuilds them each call, head.py:1226) -- also keeps host->device copies out of
        CUDA-graph capture."""
        key = (tuple(map(tuple, shapes)), dtype, str(device))
        cache = self.__dict__.setdefault("_anchor_cache", {})
        if key not in cache:
            cache.clear()
            cache[key] = self._generate_anchors(shapes, dtype=dtype, device=device)
        return cache[key]

    # Query selection (head.py:1220-1264).  The reference runs enc_output / enc_score_head over ALL B*Lv tokens with
    # autograd on, although only the B*nq selected rows ever receive a gradient (both consumers are row gathers):
    # its backward is a dense LayerNorm + two dense GEMMs (2 x 282 GFLOP at TAMTR.yaml shapes) over tensors that are
    # zero outside 0.3 % of their rows, plus two materialised [B, Lv, d] zero tensors.  Here the dense pass only ranks
    # the tokens (no graph, nothing saved); the selected rows are then recomputed differentiably, and their gradient
    # reaches `feats` as a row-sparse in-place update (ops.GradHub).  Same values, same gradients.
    sparse_query_selection = True

    def _rank_tokens(self, feats, valid):
        if getattr(feats, "is_folded", False):
            return feats.rank(self._valid_u8(valid))
        if feats.is_cuda and feats.dtype in (torch.float32, torch.bfloat16) and self.fused_input_proj:
            return ops.rank_tokens(feats, self._valid_u8(valid), self.enc_output[0], self.enc_output[1],
                                   self.enc_score_head)
        with torch.no_grad():
            features = self.enc_output(valid * feats)
            return self.enc_score_head(features).max(-1).values              # [B, Lv]

    def _valid_u8(self, valid):
        cache = self.__dict__.setdefault("_anchor_cache", {})
        key = ("valid_u8", valid.data_ptr())
        if key not in cache:
            cache[key] = valid.view(-1).to(torch.uint8).contiguous()
        return cache[key]

    def _get_decoder_input(self, feats, shapes, dn_embed=None, dn_bbox=None, hub=None):
        bs, n_tok = feats.shape[0], feats.shape[1]
        anchors, valid = self._anchors(shapes, feats.dtype, feats.device)
        if getattr(feats, "is_folded", False):
            # fold.FoldedTokens: ranking embedding and scores came out of the projection kernel; the selected rows are
            # recomputed from the NCHW maps through the folded affine map (differentiable)
            topk = ops.topk_rows(self._rank_tokens(feats, valid), self.num_queries).view(-1)
            if feats.arena is not None:
                feats.arena.start_prefill()     # the gradient arena's zero fill, forked behind the top-k kernel
            img = torch.arange(end=bs, dtype=topk.dtype, device=topk.device).unsqueeze(-1).repeat(
                1, self.num_queries).view(-1)
            rows = valid.view(-1)[topk].unsqueeze(-1) * feats.rows(img * n_tok + topk)
            top_feats = self.enc_output(rows).view(bs, self.num_queries, -1)
            enc_scores = self.enc_score_head(top_feats)
        elif self.sparse_query_selection and hub is not None:
            topk = ops.topk_rows(self._rank_tokens(feats, valid), self.num_queries).view(-1)
            img = torch.arange(end=bs, dtype=topk.dtype, device=topk.device).unsqueeze(-1).repeat(
                1, self.num_queries).view(-1)
            rows = ops.select_rows(feats, img * n_tok + topk, hub)              # [B*nq, d]
            rows = valid.view(-1)[topk].unsqueeze(-1) * rows
            top_feats = self.enc_output(rows).view(bs, self.num_queries, -1)
            enc_scores = self.enc_score_head(top_feats)
        else:
            features = self.enc_output(valid * feats)
            scores = self.enc_score_head(features)
            topk = ops.topk_rows(scores.max(-1).values, self.num_queries).view(-1)
            img = torch.arange(end=bs, dtype=topk.dtype, device=topk.device).unsqueeze(-1).repeat(
                1, self.num_queries).view(-1)
            top_feats = features[img, topk].view(bs, self.num_queries, -1)
            enc_scores = scores[img, topk].view(bs, self.num_queries, -1)
        top_anchors = anchors[:, topk].view(bs, self.num_queries, -1)
        refer_bbox = self.enc_bbox_head(top_feats) + top_anchors
        enc_bboxes = refer_bbox.sigmoid()
        if dn_bbox is not None:
            refer_bbox = torch.cat([dn_bbox, refer_bbox], 1)
        embeddings = self.tgt_embed.weight.unsqueeze(0).repeat(bs, 1, 1) if self.learnt_init_query else top_feats
        if self.training:
            refer_bbox = refer_bbox.detach()
            if not self.learnt_init_query:
                embeddings = embeddings.detach()
        if dn_embed is not None:
            embeddings = torch.cat([dn_embed, embeddings], 1)
        return embeddings, refer_bbox, enc_bboxes, enc_scores

    # Folded encoder side (fold.py): input_proj + BatchNorm + value_proj of every layer + enc_output.0 / score head as one
    # projection per level from the NCHW maps; `feats` is never materialised.  bf16 activations only.
    folded_projection = os.environ.get("TAMTR_FOLD", "1") != "0"

    def _fold_attns(self, x):
        """The decoder layers' cross-attention modules when the folded path applies to this call, else None."""
        if not (self.folded_projection and self.sparse_query_selection and self._fusable_input_proj(x)):
            return None
        lp = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x[0].dtype
        dec = self.decoder
        if lp != torch.bfloat16 or not getattr(dec, "batched_value_projection", False):
            return None
        n_used = dec.num_layers if self.training else dec.eval_idx + 1
        attns = [getattr(l, "cross_attn", None) for l in dec.layers[:n_used]]
        if not all(is_plain_msda(a) for a in attns) or any(a.n_heads != attns[0].n_heads for a in attns):
            return None
        d = self.hidden_dim
        n_tail = (self.enc_score_head.weight.shape[0] + 1 + 15) // 16 * 16
        if not (isinstance(self.enc_output, nn.Sequential) and isinstance(self.enc_output[0], nn.Linear)
                and isinstance(self.enc_output[1], nn.LayerNorm) and isinstance(self.enc_score_head, nn.Linear)
                and d <= 4 * 32 * 8 and all(a.value_proj.weight.shape == (d, d) for a in attns)):
            return None
        return attns if fold.supported(list(x), d, n_used * d, n_tail) else None

    def _encode(self, x):
        """input projection + (when the CUDA path applies) the gradient hub on the token tensor."""
        attns = self._fold_attns(x)
        if attns is not None:
            xs = [f if f.dtype == torch.bfloat16 else f.to(torch.bfloat16) for f in x]
            feats = fold.FoldedTokens(xs, self.input_proj, self.training)
            _, valid = self._anchors(feats.shapes, feats.dtype, feats.device)
            feats.project(attns, self.enc_output[0], self.enc_output[1], self.enc_score_head, self._valid_u8(valid))
            return feats, feats.shapes, None
        feats, shapes = self._get_encoder_input(x)
        hub = None
        if self.sparse_query_selection and feats.is_cuda:
            hub = ops.GradHub()
            feats = ops.grad_hub(feats, hub)
        return feats, shapes, hub

    def _cdn(self, batch):
        if isinstance(batch, DeviceTargets):   # ground truth in fixed-shape device tensors: the group is built by a kernel
            if not self.training or self.num_denoising <= 0:
                return None, None, None, None
            return cdn_group_device(batch, self.nc, self.num_queries, self.denoising_class_embed.weight,
                                    self.num_denoising, self.label_noise_ratio, self.box_noise_scale)
        if isinstance(batch, CdnPlan):      # pre-planned on the host: only the embedding gather
            return batch.materialize(self.denoising_class_embed.weight)
        return get_cdn_group(batch, self.nc, self.num_queries, self.denoising_class_embed.weight,
                             self.num_denoising, self.label_noise_ratio, self.box_noise_scale, self.training)

    def plan_cdn(self, batch):
        """Host-side half of the denoising group for `batch`; pass the result as `batch` to forward()."""
        return plan_cdn_group(batch, self.nc, self.num_queries, self.num_denoising, self.label_noise_ratio,
                              self.box_noise_scale, self.training)

    def __getstate__(self):     # keep the anchor cache out of pickles / deep copies (checkpoints, EMA)
        state = dict(self.__dict__)
        state.pop("_anchor_cache", None)
        return state

    def _finish(self, dec_bboxes, dec_scores, enc_bboxes, enc_scores, dn_meta):
        x = dec_bboxes, dec_scores, enc_bboxes, enc_scores, dn_meta
        if self.training:
            return x
        y = torch.cat((dec_bboxes.squeeze(0), dec_scores.squeeze(0).sigmoid()), -1)
        return y if self.export else (y, x)

    def _reset_common(self):
        bias_cls = float(-math.log((1 - 0.01) / 0.01)) / 80 * self.nc
        nn.init.constant_(self.enc_score_head.bias, bias_cls)
        nn.init.zeros_(self.enc_bbox_head.layers[-1].weight)
        nn.init.zeros_(self.enc_bbox_head.layers[-1].bias)
        for reg in self.dec_bbox_head:
            nn.init.zeros_(reg.layers[-1].weight)
            nn.init.zeros_(reg.layers[-1].bias)
        nn.init.xavier_uniform_(self.enc_output[0].weight)
        bound = 1 / math.sqrt(self.enc_output[0].weight.shape[0])
        nn.init.uniform_(self.enc_output[0].bias, -bound, bound)
        if self.learnt_init_query:
            nn.init.xavier_uniform_(self.tgt_embed.weight)
        nn.init.xavier_uniform_(self.query_pos_head.layers[0].weight)
        nn.init.xavier_uniform_(self.query_pos_head.layers[1].weight)
        for layer in self.input_proj:
            nn.init.xavier_uniform_(layer[0].weight)
        return bias_cls


class RTDETRDecoder(_HeadBase):
    """RT-DETR head (head.py:174-435): forward(x: list of NCHW maps, batch=None)."""

    def __init__(self, nc=80, ch=(512, 1024, 2048), hd=256, nq=300, ndp=4, nh=8, ndl=6, d_ffn=1024, eval_idx=-1,
                 dropout=0., act=nn.ReLU(), nd=100, label_noise_ratio=0.5, box_noise_scale=1.0,
                 learnt_init_query=False):
        super().__init__()
        self.nhead = nh
        self.num_decoder_layers = ndl
        self._build_common(nc, ch, hd, nq, nd, label_noise_ratio, box_noise_scale, learnt_init_query)
        layer = DeformableTransformerDecoderLayer(hd, nh, d_ffn, dropout, act, self.nl, ndp)
        self.decoder = DeformableTransformerDecoder(hd, layer, ndl, eval_idx)
        self.dec_score_head = nn.ModuleList([nn.Linear(hd, nc) for _ in range(ndl)])
        self.dec_bbox_head = nn.ModuleList([MLP(hd, hd, 4, num_layers=3) for _ in range(ndl)])
        bias_cls = self._reset_common()
        for cls in self.dec_score_head:
            nn.init.constant_(cls.bias, bias_cls)

    def forward(self, x, batch=None):
        feats, shapes, hub = self._encode(x)
        dn_embed, dn_bbox, attn_mask, dn_meta = self._cdn(batch)
        embed, refer_bbox, enc_bboxes, enc_scores = self._get_decoder_input(feats, shapes, dn_embed, dn_bbox, hub)
        dec_bboxes, dec_scores = self.decoder(embed, refer_bbox, feats, shapes, self.dec_bbox_head,
                                              self.dec_score_head, self.query_pos_head, attn_mask=attn_mask)
        return self._finish(dec_bboxes, dec_scores, enc_bboxes, enc_scores, dn_meta)


VSS_PARALLEL_LEVELS = os.environ.get("TAMTR_VSS_PARALLEL", "1") != "0"
_VSS_STREAMS = {}


def _apply_vss_blocks(blocks, x):
    """One VSSBlock per pyramid level (head.py:1134).  The levels are independent, and the selective scan of the largest
    level runs with two warps per scheduler (16 384 channels): each level is issued on its own stream -- forked from and
    joined to the caller's stream, so it is a parallel branch of a captured step, and autograd runs each block's backward on
    its forward stream -- so that the smaller levels' kernels fill the SMs the big scan leaves idle."""
    if not (VSS_PARALLEL_LEVELS and x[0].is_cuda and len(x) > 1 and not isinstance(blocks[0], nn.Identity)):
        return [blk(f.permute(0, 2, 3, 1)).permute(0, 3, 1, 2) for blk, f in zip(blocks, x)]
    dev = x[0].device
    side = _VSS_STREAMS.get((dev, len(x)))
    if side is None:
        side = _VSS_STREAMS[(dev, len(x))] = [torch.cuda.Stream(dev) for _ in range(len(x) - 1)]
    main = torch.cuda.current_stream(dev)
    outs = [None] * len(x)
    for i in range(1, len(x)):                                    # levels 1.. on side streams, level 0 on the caller's
        side[i - 1].wait_stream(main)
        with torch.cuda.stream(side[i - 1]):
            outs[i] = blocks[i](x[i].permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    outs[0] = blocks[0](x[0].permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    for s in side:
        main.wait_stream(s)
    return outs


class ManbaWorldDecoder(_HeadBase):
    """TAM-TR's Multi-modal Encoder-decoder Head (head.py:1005-1290): forward(x, text [B,K,512], batch=None).

    `VSSBlocks` (head.py:1092-1098): one VMamba VSSBlock per pyramid level in front of input_proj, on the selective-scan
    kernels of csrc/sscan.cu (tamtr_b200/vss.py).  `vss=False` replaces them by identities -- the configuration the
    head-level parity fixtures and the CPU reference arm use, because the reference cannot run its own VSSBlocks without
    the un-vendored `selective_scan_cuda_core` extension; `VSSBlocks.*` checkpoint entries are then skipped on load."""

    def __init__(self, nc=80, ch=(512, 1024, 2048), hd=512, nq=300, ndp=4, nh=8, ndl=6, d_ffn=1024, eval_idx=-1,
                 dropout=0., act=nn.ReLU(), nd=100, label_noise_ratio=0.5, box_noise_scale=1.0,
                 learnt_init_query=False, dims=(128, 256, 512), drop_path=(0.1, 0.1, 0.1), embed=512, with_bn=False,
                 vss=True):
        super().__init__()
        if with_bn:
            raise NotImplementedError("tamtr_b200: BNContrastiveHeadMLP (with_bn=True) is not on the TAMTR.yaml path")
        self.nhead = nh
        self.num_decoder_layers = ndl
        self._build_common(nc, ch, hd, nq, nd, label_noise_ratio, box_noise_scale, learnt_init_query)
        self.vss = bool(vss)
        self.VSSBlocks = nn.ModuleList(VSSBlock(hidden_dim=d, drop_path=p) if vss else nn.Identity()
                                       for d, p in zip(dims, drop_path))
        self.num_Blocks = len(dims)
        layer = DeformableTransformerDecoderLayer(hd, nh, d_ffn, dropout, act, self.nl, ndp)
        self.decoder = TextDeformableTransformerDecoder(hd, layer, ndl, eval_idx)
        self.dec_score_head = nn.ModuleList([ContrastiveHeadMLP() for _ in range(ndl)])
        self.dec_bbox_head = nn.ModuleList([MLP(hd, hd, 4, num_layers=3) for _ in range(ndl)])
        self._reset_common()

    def forward(self, x, text, batch=None):
        if getattr(self, "vss", True):        # head.py:1134: channel-last in and out (reference instances: always)
            x = _apply_vss_blocks(self.VSSBlocks, x)
        feats, shapes, hub = self._encode(x)
        dn_embed, dn_bbox, attn_mask, dn_meta = self._cdn(batch)
        embed, refer_bbox, enc_bboxes, enc_scores = self._get_decoder_input(feats, shapes, dn_embed, dn_bbox, hub)
        dec_bboxes, dec_scores = self.decoder(embed, refer_bbox, feats, shapes, text, self.dec_bbox_head,
                                              self.dec_score_head, self.query_pos_head, attn_mask=attn_mask)
        return self._finish(dec_bboxes, dec_scores, enc_bboxes, enc_scores, dn_meta)

    def load_state_dict(self, state_dict, strict=True, **kw):
        if not self.vss:
            state_dict = {k: v for k, v in state_dict.items() if not k.startswith("VSSBlocks.")}
        return super().load_state_dict(state_dict, strict=strict, **kw)
