"""Host-side mirror of the reference's module interface for the detection-head hot path.

Same class names, constructor arguments, forward signatures, attribute / parameter names (state_dict keys) and
error behaviour as the reference, so its checkpoints load and its callers (`parse_model`, `RTDETRDecoder`,
`ManbaWorldDecoder`) can use these classes unchanged -- but the arithmetic the north star names runs in the
hand-written sm_100a kernels behind include/tamtr_b200.h (ops.py).  Everything else (GEMMs, LayerNorm, the small
query self-attention) is a library call, as in the reference.

Reference classes mirrored (file:line under /root/reference/ultralytics):
  MLP                                  nn/modules/transformer.py:162-176
  MSDeformAttn                         nn/modules/transformer.py:204-299
  DeformableTransformerDecoderLayer    nn/modules/transformer.py:498-558
  DeformableTransformerDecoder         nn/modules/transformer.py:662-716
  TextDeformableTransformerDecoder     nn/modules/transformer.py:835-891
  ContrastiveHeadMLP                   nn/modules/block.py:522-541
  MaxSigmoidAttnBlock                  nn/extra_modules/block.py:194-226
  TIAGELAN                             nn/extra_modules/block.py:171-192
  inverse_sigmoid                      nn/modules/utils.py:34-39
"""
import copy
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

__all__ = ("MLP", "MSDeformAttn", "MSDeformAttncls", "MSDeformAttnbox", "DeformableTransformerDecoderLayer",
           "DecouplingDeformableTransformerDecoderLayer", "DeformableTransformerDecoder",
           "TextDeformableTransformerDecoder", "ContrastiveHeadMLP", "MaxSigmoidAttnBlock", "TIAGELAN",
           "inverse_sigmoid")


def inverse_sigmoid(x, eps=1e-5):
    """log(x / (1 - x)) with both terms clamped to eps (utils.py:34-39)."""
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def _clones(module, n):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(n)])


class MLP(nn.Module):
    """Linear -> ReLU -> ... -> Linear (transformer.py:162-176); parameters live under `layers.{i}`."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x):
        last = self.num_layers - 1
        for i, layer in enumerate(self.layers):
            x = ops.linear(x, layer) if i == last else ops.linear_relu(x, layer)
        return x


class MSDeformAttn(nn.Module):
    """Multi-scale deformable attention (transformer.py:204-299) on the sm_100a sampler.

    forward(query [B,Lq,C], refer_bbox [B,Lq,n_levels|1,2|4], value [B,Lv,C], value_shapes [[h,w]]*L,
            value_mask [B,Lv] | None) -> [B,Lq,C]
    """

    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError(f"d_model must be divisible by n_heads, but got {d_model} and {n_heads}")
        self.im2col_step = 64
        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        # transformer.py:234-250: offsets start as a ring of directions (one per head) scaled by the point index,
        # attention logits start at zero (uniform 1/(L*P) weights), projections are Xavier.
        nn.init.zeros_(self.sampling_offsets.weight)
        theta = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        ring = torch.stack([theta.cos(), theta.sin()], -1)
        ring = ring / ring.abs().max(-1, keepdim=True)[0]
        ring = ring.view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        ring = ring * torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(ring.reshape(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.zeros_(self.value_proj.bias)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.zeros_(self.output_proj.bias)

    def forward(self, query, refer_bbox, value, value_shapes, value_mask=None, projected_value=None, arena=None):
        """`projected_value` / `arena`: optional [B,Lv,H,Dh] view produced by the decoder's batched value projection
        (ops.split_values) -- then `value` is not projected again here."""
        bs, len_q = query.shape[:2]
        len_v = value.shape[1]
        assert sum(s[0] * s[1] for s in value_shapes) == len_v
        if projected_value is None:
            value = self.value_proj(value)
            if value_mask is not None:
                value = value.masked_fill(value_mask[..., None], float(0))
            value = value.view(bs, len_v, self.n_heads, self.d_model // self.n_heads)
        else:
            value = projected_value
        loc, attn = ops.sampling_locations_and_weights(
            query, refer_bbox, self.sampling_offsets.weight, self.sampling_offsets.bias,
            self.attention_weights.weight, self.attention_weights.bias, value_shapes,
            self.n_heads, self.n_levels, self.n_points)
        out = ops.ms_deform_attn(value, value_shapes, loc, attn, arena)
        return ops.linear(out, self.output_proj)


def _ragged_attn_forward(attn_mod, points, query, refer_bbox, value, value_shapes, value_mask):
    """MSDeformAttncls / MSDeformAttnbox forward (transformer.py:347-397, :445-495): the projections, softmax and
    location math are those of MSDeformAttn on the flat [.., n_levels*n_points, 2] view; only the split of the samples
    over the levels differs (`points`, utils.py:108 / :159)."""
    bs, len_q = query.shape[:2]
    len_v = value.shape[1]
    assert sum(s[0] * s[1] for s in value_shapes) == len_v
    n_s = attn_mod.n_levels * attn_mod.n_points
    if len(value_shapes) != len(points) or n_s != sum(points):
        raise RuntimeError(f"tamtr_b200: the reference splits {n_s} samples as {list(points)} over {len(points)} levels "
                           f"(utils.py:108,159); got n_levels={attn_mod.n_levels}, n_points={attn_mod.n_points}, "
                           f"{len(value_shapes)} value shapes")
    if refer_bbox.shape[-1] != 4:
        # the reference's 2-d branch broadcasts a 5-d offset tensor against a 6-d normaliser and cannot run
        raise ValueError(f"Last dim of reference_points must be 2 or 4, but got {refer_bbox.shape[-1]}.")
    value = attn_mod.value_proj(value)
    if value_mask is not None:
        value = value.masked_fill(value_mask[..., None], float(0))
    value = value.view(bs, len_v, attn_mod.n_heads, attn_mod.d_model // attn_mod.n_heads)
    loc, attn = ops.sampling_locations_and_weights(
        query, refer_bbox, attn_mod.sampling_offsets.weight, attn_mod.sampling_offsets.bias,
        attn_mod.attention_weights.weight, attn_mod.attention_weights.bias, value_shapes,
        attn_mod.n_heads, attn_mod.n_levels, attn_mod.n_points)
    out = ops.ms_deform_attn_ragged(value, value_shapes, loc.view(bs, len_q, attn_mod.n_heads, n_s, 2), attn, points)
    return ops.linear(out, attn_mod.output_proj)


class MSDeformAttncls(MSDeformAttn):
    """Classification-branch deformable attention of the decoupled decoder layer (transformer.py:300-397):
    2 / 4 / 6 points on the three pyramid levels, largest map first."""

    def forward(self, query, refer_bbox, value, value_shapes, value_mask=None):
        return _ragged_attn_forward(self, (2, 4, 6), query, refer_bbox, value, value_shapes, value_mask)


class MSDeformAttnbox(MSDeformAttn):
    """Box-branch deformable attention (transformer.py:400-495): 6 / 4 / 2 points on the three levels."""

    def forward(self, query, refer_bbox, value, value_shapes, value_mask=None):
        return _ragged_attn_forward(self, (6, 4, 2), query, refer_bbox, value, value_shapes, value_mask)


# The forward() methods below are also bound onto the REFERENCE's classes by patch.enable(), so they may only touch
# attributes the reference's __init__ creates: helpers are module-level functions, not methods.
def _add_norm(layer, x, y, drop, norm):
    """norm(x + dropout(y)): fused residual + LayerNorm kernel (dropout is the identity in TAM-TR, p = 0)."""
    if ((drop.p == 0.0 or not layer.training) and ops.add_layer_norm_supported(x, x.shape[-1])
            and isinstance(norm, nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and tuple(norm.normalized_shape) == (x.shape[-1],)):
        return ops.add_layer_norm(x, y, norm)
    return norm(x + drop(y))


def _ffn(layer, tgt):
    if isinstance(layer.act, nn.ReLU):
        hidden = ops.linear_relu(tgt, layer.linear1)
    else:
        hidden = layer.act(ops.linear(tgt, layer.linear1))
    tgt2 = ops.linear(layer.dropout3(hidden), layer.linear2)
    return _add_norm(layer, tgt, tgt2, layer.dropout4, layer.norm3)


LOWP_LAYER = os.environ.get("TAMTR_LOWP_LAYER", "1") != "0"


def _plain_norm(layer, drop, norm, d):
    return ((drop.p == 0.0 or not layer.training) and isinstance(norm, nn.LayerNorm) and norm.elementwise_affine
            and norm.bias is not None and tuple(norm.normalized_shape) == (d,))


def _layer_lowp_applies(layer, embed, query_pos):
    """bf16 autocast over an fp32 query stream: the case where autocast surrounds every projection of the layer with casts
    (and the `+ query_pos` adds run in fp32) -- those become side outputs of the add + LayerNorm kernels."""
    if not (LOWP_LAYER and embed.is_cuda and embed.dtype == torch.float32 and query_pos is not None
            and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16):
        return False
    d = embed.shape[-1]
    mha = layer.self_attn
    return (query_pos.shape == embed.shape and query_pos.dtype in (torch.float32, torch.bfloat16)
            and embed.numel() % 4 == 0 and ops.add_layer_norm_supported(embed, d)
            and mha._qkv_same_embed_dim and mha.in_proj_bias is not None and mha.bias_k is None and not mha.add_zero_attn
            and _plain_norm(layer, layer.dropout1, layer.norm1, d) and _plain_norm(layer, layer.dropout2, layer.norm2, d)
            and _plain_norm(layer, layer.dropout4, layer.norm3, d) and (layer.dropout3.p == 0.0 or not layer.training))


def _layer_forward_lowp(layer, embed, refer_bbox, feats, shapes, padding_mask, attn_mask, query_pos, projected_value, arena):
    """DeformableTransformerDecoderLayer.forward (transformer.py:535-558) with the same arithmetic as the autocast path --
    fp32 stream, bf16 operands rounded from the same fp32 values -- in 1 + 3 element-wise launches per layer instead of
    3 + 6: bf16(embed) / bf16(embed + pos) come from one pass over the layer's input, bf16(norm1(..) + pos) and
    bf16(norm2(..)) from the add + LayerNorm kernels, and the backward sums the gradients of those copies inside the
    kernels that consume them."""
    embed, e_lp, q_lp = ops.pos_cast(embed, query_pos)
    tgt = ops.self_attention(layer.self_attn, q_lp, e_lp, attn_mask)
    embed, _, q_lp = ops.add_layer_norm_sides(embed, tgt, layer.norm1, pos=query_pos)
    tgt = layer.cross_attn(q_lp, refer_bbox.unsqueeze(2), feats, shapes, padding_mask, projected_value, arena)
    embed, e_lp, _ = ops.add_layer_norm_sides(embed, tgt, layer.norm2, want_lp=True)
    if isinstance(layer.act, nn.ReLU):
        hidden = ops.linear_relu(e_lp, layer.linear1)
    else:
        hidden = layer.act(ops.linear(e_lp, layer.linear1))
    tgt2 = ops.linear(layer.dropout3(hidden), layer.linear2)
    return ops.add_layer_norm(embed, tgt2, layer.norm3)


def _folded_bn(block, bn):
    """BatchNorm2d in eval mode as a per-channel affine (fp32), cached on the block until a parameter or statistic
    changes (the entry keeps the four source tensors and is valid only for those very objects at those versions)."""
    src = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    hit = block.__dict__.get("_tamtr_fold_cache")
    if hit is None or any(a is not b for a, b in zip(hit[0], src)) or hit[1] != tuple(t._version for t in src):
        s = bn.weight.detach().float() * torch.rsqrt(bn.running_var.float() + bn.eps)
        t = bn.bias.detach().float() - bn.running_mean.float() * s
        hit = (src, tuple(t_._version for t_ in src), s.contiguous(), t.contiguous())
        block.__dict__["_tamtr_fold_cache"] = hit
    return hit[2], hit[3]


class DeformableTransformerDecoderLayer(nn.Module):
    """Self-attention -> deformable cross-attention -> FFN, post-norm (transformer.py:498-558)."""

    def __init__(self, d_model=256, n_heads=8, d_ffn=1024, dropout=0., act=nn.ReLU(), n_levels=4, n_points=4):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.cross_attn = MSDeformAttn(d_model, n_levels, n_heads, n_points)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.act = act
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn(self, tgt):
        return _ffn(self, tgt)

    def forward(self, embed, refer_bbox, feats, shapes, padding_mask=None, attn_mask=None, query_pos=None,
                projected_value=None, arena=None):
        if _layer_lowp_applies(self, embed, query_pos):
            return _layer_forward_lowp(self, embed, refer_bbox, feats, shapes, padding_mask, attn_mask, query_pos,
                                       projected_value, arena)
        q = k = self.with_pos_embed(embed, query_pos)
        # need_weights=False: the reference discards the averaged attention map ([0] only, transformer.py:546-547),
        # so the fused SDPA kernels can be used; the output is the same.
        mha = self.self_attn
        if (embed.is_cuda and mha._qkv_same_embed_dim and mha.in_proj_bias is not None and mha.bias_k is None
                and not mha.add_zero_attn):
            tgt = ops.self_attention(mha, q, embed, attn_mask)
        else:
            tgt = mha(q.transpose(0, 1), k.transpose(0, 1), embed.transpose(0, 1), attn_mask=attn_mask,
                      need_weights=False)[0].transpose(0, 1)
        embed = _add_norm(self, embed, tgt, self.dropout1, self.norm1)
        tgt = self.cross_attn(self.with_pos_embed(embed, query_pos), refer_bbox.unsqueeze(2), feats, shapes,
                              padding_mask, projected_value, arena)
        embed = _add_norm(self, embed, tgt, self.dropout2, self.norm2)
        return _ffn(self, embed)


def _self_attention(mha, q, k, v, attn_mask):
    if (v.is_cuda and mha._qkv_same_embed_dim and mha.in_proj_bias is not None and mha.bias_k is None
            and not mha.add_zero_attn):
        return ops.self_attention(mha, q, v, attn_mask)
    return mha(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), attn_mask=attn_mask,
               need_weights=False)[0].transpose(0, 1)


class DecouplingDeformableTransformerDecoderLayer(nn.Module):
    """Decoder layer with separate classification / box streams (transformer.py:561-658): self-attention on the
    classification stream, MSDeformAttncls on it, MSDeformAttnbox on the box stream, one FFN each."""

    def __init__(self, d_model=256, n_heads=8, d_ffn=1024, dropout=0., act=nn.ReLU(), n_levels=4, n_points=4):
        super().__init__()
        self.self_attn1 = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.cross_attn_cls = MSDeformAttncls(d_model, n_levels, n_heads, n_points)
        self.cross_attn_box = MSDeformAttnbox(d_model, n_levels, n_heads, n_points)
        self.dropout3 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm4 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.act = act
        self.dropout5 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout6 = nn.Dropout(dropout)
        self.norm5 = nn.LayerNorm(d_model)
        self.linear3 = nn.Linear(d_model, d_ffn)
        self.dropout7 = nn.Dropout(dropout)
        self.linear4 = nn.Linear(d_ffn, d_model)
        self.dropout8 = nn.Dropout(dropout)
        self.norm6 = nn.LayerNorm(d_model)

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def forward_ffn1(self, tgt):
        tgt2 = ops.linear(self.dropout5(self.act(ops.linear(tgt, self.linear1))), self.linear2)
        return _add_norm(self, tgt, tgt2, self.dropout6, self.norm5)

    def forward_ffn2(self, tgt):
        tgt2 = ops.linear(self.dropout7(self.act(ops.linear(tgt, self.linear3))), self.linear4)
        return _add_norm(self, tgt, tgt2, self.dropout8, self.norm6)

    def forward(self, embed, embed1, refer_bbox, feats, shapes, padding_mask=None, attn_mask=None, query_pos=None,
                dn_meta=None):
        with_pos = DecouplingDeformableTransformerDecoderLayer.with_pos_embed
        q = k = with_pos(embed, query_pos)
        tgt = _self_attention(self.self_attn1, q, k, embed, attn_mask)
        embed = _add_norm(self, embed, tgt, self.dropout1, self.norm1)
        tgt = self.cross_attn_cls(with_pos(embed, query_pos), refer_bbox.unsqueeze(2), feats, shapes, padding_mask)
        embed = _add_norm(self, embed, tgt, self.dropout3, self.norm3)
        tgt = self.cross_attn_box(with_pos(embed1, query_pos), refer_bbox.unsqueeze(2), feats, shapes, padding_mask)
        embed1 = _add_norm(self, embed1, tgt, self.dropout4, self.norm4)
        # forward_ffn1 / forward_ffn2 (transformer.py:609-619), written out so that the patched reference class needs no
        # other rebinding
        t2 = ops.linear(self.dropout5(self.act(ops.linear(embed, self.linear1))), self.linear2)
        out_cls = _add_norm(self, embed, t2, self.dropout6, self.norm5)
        t2 = ops.linear(self.dropout7(self.act(ops.linear(embed1, self.linear3))), self.linear4)
        return out_cls, _add_norm(self, embed1, t2, self.dropout8, self.norm6)


def is_plain_msda(a):
    """MSDeformAttn itself (ours or the reference's class of that name, whose forward enable() rebinds) -- not the
    ragged _cls / _box variants, whose value tensors are consumed differently."""
    return a is not None and type(a).__name__ == "MSDeformAttn" and isinstance(getattr(a, "value_proj", None), nn.Linear)


class _DecoderBase(nn.Module):
    def __init__(self, hidden_dim, decoder_layer, num_layers, eval_idx=-1):
        super().__init__()
        self.layers = _clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.hidden_dim = hidden_dim
        self.eval_idx = eval_idx if eval_idx >= 0 else num_layers + eval_idx

    batched_value_projection = True

    def _project_values(self, feats, padding_mask, n_used):
        """transformer.py:273 runs value_proj inside every layer on the SAME `feats` (transformer.py:870): do all
        layers with one [d, n*d] GEMM and hand each layer a column-slice view (head-major, what the sampler reads)."""
        if getattr(feats, "is_folded", False):      # fold.FoldedTokens: the head projected them with the folded weights
            return feats.values, feats.arena
        if not (self.batched_value_projection and feats.is_cuda and padding_mask is None and n_used > 1):
            return [None] * n_used, None
        attns = [getattr(l, "cross_attn", None) for l in self.layers[:n_used]]
        if not all(is_plain_msda(a) for a in attns):
            return [None] * n_used, None
        w = torch.cat([a.value_proj.weight for a in attns], 0)
        b = torch.cat([a.value_proj.bias for a in attns], 0)
        if not (torch.is_grad_enabled() and (feats.requires_grad or w.requires_grad)):
            value_all = F.linear(feats, w, b)
            d = attns[0].d_model
            bs, lv = feats.shape[:2]
            return [value_all[:, :, i * d:(i + 1) * d].view(bs, lv, attns[0].n_heads, -1) for i in range(n_used)], None
        arena = ops.ValueArena()
        return list(ops.project_values(feats, w, b, arena, n_used, attns[0].n_heads)), arena

    def _run(self, embed, refer_bbox, feats, shapes, bbox_head, score_fn, pos_mlp, attn_mask, padding_mask):
        output = embed
        dec_bboxes, dec_cls = [], []
        last_refined = None
        refer_bbox = refer_bbox.sigmoid()
        n_used = self.num_layers if self.training else self.eval_idx + 1
        values, arena = self._project_values(feats, padding_mask, n_used)
        for i, layer in enumerate(self.layers):
            output = layer(output, refer_bbox, feats, shapes, padding_mask, attn_mask, pos_mlp(refer_bbox),
                           values[i] if i < n_used else None, arena)
            bbox = bbox_head[i](output)
            refine = ops.box_refine          # sigmoid(bbox + inverse_sigmoid(ref)) as one kernel
            refined = refine(bbox, refer_bbox)
            if self.training:
                dec_cls.append(score_fn(i, output))
                dec_bboxes.append(refined if i == 0 else refine(bbox, last_refined))
            elif i == self.eval_idx:
                dec_cls.append(score_fn(i, output))
                dec_bboxes.append(refined)
                break
            last_refined = refined
            refer_bbox = refined.detach() if self.training else refined
        return torch.stack(dec_bboxes), torch.stack(dec_cls)


class DeformableTransformerDecoder(_DecoderBase):
    """transformer.py:662-716; score heads are plain Linear layers."""

    def forward(self, embed, refer_bbox, feats, shapes, bbox_head, score_head, pos_mlp, attn_mask=None,
                padding_mask=None):
        return self._run(embed, refer_bbox, feats, shapes, bbox_head, lambda i, x: score_head[i](x), pos_mlp,
                         attn_mask, padding_mask)


class TextDeformableTransformerDecoder(_DecoderBase):
    """transformer.py:835-891; score heads take the text embeddings (ContrastiveHeadMLP)."""

    def forward(self, embed, refer_bbox, feats, shapes, text, bbox_head, score_head, pos_mlp, attn_mask=None,
                padding_mask=None):
        return self._run(embed, refer_bbox, feats, shapes, bbox_head, lambda i, x: score_head[i](x, text), pos_mlp,
                         attn_mask, padding_mask)


class ContrastiveHeadMLP(nn.Module):
    """Region-text similarity (block.py:522-541): cosine(x[b,q,:], w[b,k,:]) * exp(logit_scale) + bias."""

    def __init__(self):
        super().__init__()
        self.bias = nn.Parameter(torch.tensor([-10.0]))
        self.logit_scale = nn.Parameter(torch.ones([]) * torch.tensor(1 / 0.07).log())

    def forward(self, x, w):
        return ops.contrastive_head(x, w, self.logit_scale, self.bias)


class _ConvBN(nn.Module):
    """ultralytics Conv(c1, c2, k, act=...) = Conv2d(bias=False) + BatchNorm2d (+ SiLU), keys `conv.*` / `bn.*`
    (nn/modules/conv.py:23-40)."""

    def __init__(self, c1, c2, k, act=False):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, 1, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class MaxSigmoidAttnBlock(nn.Module):
    """BTA-PAN text-image attention (extra_modules/block.py:194-226): per head, the max over text tokens of
    <x, gl(guide)> / sqrt(hc) + bias, through a sigmoid, gates a 3x3 conv of x."""

    def __init__(self, c1, c2, nh=1, ec=128, gc=512, scale=False):
        super().__init__()
        self.nh = nh
        self.hc = c2 // nh
        self.ec = _ConvBN(c1, ec, 1) if c1 != ec else None
        self.gl = nn.Linear(gc, ec)
        self.bias = nn.Parameter(torch.zeros(nh))
        self.proj_conv = _ConvBN(c1, c2, 3)
        self.scale = nn.Parameter(torch.ones(1, nh, 1, 1)) if scale else 1.0

    def forward(self, x, guide):
        bs, _, h, w = x.shape
        guide = self.gl(guide).view(bs, -1, self.nh, self.hc)
        embed = self.ec(x) if self.ec is not None else x
        aw = ops.max_sigmoid_gate(embed, guide, self.bias, self.nh)          # [B, nh, H, W]
        aw = aw * self.scale
        pc = self.proj_conv
        conv, bn = pc.conv, getattr(pc, "bn", None)
        infer = not (self.training or torch.is_grad_enabled())
        if not ops.gate_conv3x3_supported(x, conv.weight, self.nh):
            y = pc(x)
        elif bn is None:
            # after model.fuse() (nn/tasks.py:131-136): BatchNorm folded into `conv`, which now carries a bias, and the
            # `bn` attribute deleted -> scale 1, shift = that bias
            shift = conv.bias if conv.bias is not None else torch.zeros(conv.out_channels, device=x.device)
            if infer:
                return ops.gate_conv3x3(x, conv.weight, torch.ones_like(shift, dtype=torch.float32), shift, aw, self.nh)
            y = ops.conv3x3_tc(x, conv.weight) + shift.view(1, -1, 1, 1).to(x.dtype)
        elif conv.bias is not None:
            y = pc(x)
        elif infer and bn.running_var is not None:
            # inference: conv + folded BatchNorm + gate in one tensor-core kernel (channels-last output)
            s, t = _folded_bn(self, bn)
            return ops.gate_conv3x3(x, conv.weight, s, t, aw, self.nh)
        else:
            y = bn(ops.conv3x3_tc(x, conv.weight))
        return (y.view(bs, self.nh, -1, h, w) * aw.unsqueeze(2).to(y.dtype)).view(bs, -1, h, w)


def _discarded_attn(attn, x):
    """TIAGELAN calls `self.attn(y[-3], guide)` and DROPS the result (extra_modules/block.py:185), so the text-image
    gate, the multiply and -- in eval mode -- the whole block are dead code.  What survives is the train-mode side
    effect: the BatchNorm layers inside the block (`proj_conv.bn`, and `ec.bn` when c1 != ec) see a batch and update
    their running statistics.  Reproduce exactly that, nothing else; no autograd graph (the reference builds one and
    never back-propagates it: these parameters are the "unused" ones DDP is told about)."""
    if not attn.training:
        return
    with torch.no_grad():
        for unit in (attn.ec, attn.proj_conv):
            bn = None if unit is None else getattr(unit, "bn", None)
            if bn is None or not bn.track_running_stats:
                continue
            conv = unit.conv
            if unit is attn.proj_conv and conv.bias is None and ops.gate_conv3x3_supported(x, conv.weight, 1):
                bn(ops.conv3x3_tc(x, conv.weight))
            else:
                bn(conv(x))


class _RepConvN(nn.Module):
    """Training form of the reference's RepConvN (extra_modules/block.py:24-50): SiLU(conv3x3+BN + conv1x1+BN)."""

    def __init__(self, c1, c2):
        super().__init__()
        self.conv1 = _ConvBN(c1, c2, 3)
        self.conv2 = _ConvBN(c1, c2, 1)
        self.act = nn.SiLU()

    def forward(self, x):
        return self.act(self.conv1(x) + self.conv2(x))


class _RepNBottleneck(nn.Module):
    def __init__(self, c):          # extra_modules/block.py:126-136 with c1 == c2, e = 1
        super().__init__()
        self.cv1 = _RepConvN(c, c)
        self.cv2 = _ConvBN(c, c, 3, act=True)

    def forward(self, x):
        return x + self.cv2(self.cv1(x))


class _RepNCSP(nn.Module):
    def __init__(self, c1, c2, n=1):   # extra_modules/block.py:138-149
        super().__init__()
        c_ = c2 // 2
        self.cv1 = _ConvBN(c1, c_, 1, act=True)
        self.cv2 = _ConvBN(c1, c_, 1, act=True)
        self.cv3 = _ConvBN(2 * c_, c2, 1, act=True)
        self.m = nn.Sequential(*(_RepNBottleneck(c_) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class TIAGELAN(nn.Module):
    """BTA-PAN stage (extra_modules/block.py:171-192): a CSP-ELAN block that also owns a MaxSigmoidAttnBlock whose
    output it discards.  Same submodule names / state_dict keys as the reference; the convolutions around the attention
    block are library calls (they are the neck, not this path).  forward(x [B,c1,H,W], guide [B,N,512]) -> [B,c2,H,W],
    independent of `guide` (SURVEY.md KAT #4)."""

    def __init__(self, c1, c2, c3, c4, c5=1, nh=8):
        super().__init__()
        self.c = c3 // 2
        self.cv1 = _ConvBN(c1, c3, 1, act=True)
        self.cv2 = nn.Sequential(_RepNCSP(c3 // 2, c4, c5), _ConvBN(c4, c4, 3, act=True))
        self.cv3 = nn.Sequential(_RepNCSP(c4, c4, c5), _ConvBN(c4, c4, 3, act=True))
        self.cv4 = _ConvBN(c3 + 2 * c4, c2, 1, act=True)
        self.attn = MaxSigmoidAttnBlock(c4, c4, nh=nh, ec=c4)

    def forward(self, x, guide):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in (self.cv2, self.cv3))
        _discarded_attn(self.attn, y[-3])
        return self.cv4(torch.cat(y, 1))
