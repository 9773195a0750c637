"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference in the build container.

    python -m oracle.make_goldens [names...]

/root/reference is pure PyTorch, so "outputs of the reference itself run here" are produced by importing it
(oracle/reference_loader.py) on CPU, fp32.  Inputs are regenerated from seeds by the tests (CPU generators are
machine-independent), so the fixtures only hold seeds, shapes and reference OUTPUTS (full tensors when small,
otherwise a fixed random subset + norms).  Every fixture records the reference file:line that produced it.
"""
import os
import sys

import numpy as np
import torch

from . import msda, reference_loader, seeding

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _subset(t, n, seed):
    """Fixed random subset of a big tensor: (flat indices int64, values)."""
    flat = t.detach().reshape(-1)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randperm(flat.numel(), generator=g)[:n].sort().values
    return idx.to(torch.int32), flat[idx].clone()


def _save(name, obj):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".pt")
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e3:.1f} kB)")


# ------------------------------------------------------------------------------------------------ core op
CORE_CASES = {
    # name: (seed, B, Lq, H, Dh, shapes, oob_frac)
    "tiny_nonsquare": (11, 2, 17, 4, 8, [[9, 13], [5, 7], [3, 4]], 0.30),
    "small_dh32": (12, 2, 50, 8, 32, [[20, 20], [10, 10], [5, 5]], 0.25),
    "small_dh64": (13, 2, 40, 8, 64, [[24, 16], [12, 8], [6, 4]], 0.25),
    "small_dh16_L4": (14, 1, 33, 4, 16, [[16, 16], [8, 8], [4, 4], [2, 2]], 0.25),
    "sbase_b2": (15, 2, 300, 8, 32, [[80, 80], [40, 40], [20, 20]], 0.25),       # BASELINE.json config 1 shapes
    "syaml_b1": (16, 1, 300, 8, 64, [[160, 160], [80, 80], [40, 40]], 0.25),     # TAMTR.yaml:67 shapes
}


def gen_core(ns):
    ref = ns.multi_scale_deformable_attn_pytorch   # ultralytics/nn/modules/utils.py:42
    out = {"source": "ultralytics/nn/modules/utils.py:42-89 multi_scale_deformable_attn_pytorch, torch "
                     + torch.__version__ + " CPU fp32", "cases": {}}
    for name, (seed, B, Lq, H, Dh, shapes, oob) in CORE_CASES.items():
        value, loc, attn, grad_out = msda.make_inputs(seed, B, Lq, H, Dh, shapes, oob_frac=oob)
        value.requires_grad_(), loc.requires_grad_(), attn.requires_grad_()
        o = ref(value, shapes, loc, attn)
        o.backward(grad_out)
        case = dict(seed=seed, B=B, Lq=Lq, H=H, Dh=Dh, shapes=shapes, oob_frac=oob, P=4,
                    out_norm=o.detach().double().norm().item(),
                    grad_value_norm=value.grad.double().norm().item(),
                    grad_loc_norm=loc.grad.double().norm().item(),
                    grad_attn_norm=attn.grad.double().norm().item())
        if value.numel() <= 300_000:
            case.update(out=o.detach().clone(), grad_value=value.grad.clone(), grad_loc=loc.grad.clone(),
                        grad_attn=attn.grad.clone())
        else:   # full-size shapes: fixed random subsets + norms keep the fixture small
            case.update(out_subset=_subset(o, 8000, seed + 1001), grad_value_subset=_subset(value.grad, 8000, seed + 1000),
                        grad_loc_subset=_subset(loc.grad, 8000, seed + 1002),
                        grad_attn_subset=_subset(attn.grad, 8000, seed + 1003))
        out["cases"][name] = case
    _save("msda_core", out)


def gen_index_probe(ns):
    """Bit-exact index contract: identity images expose the bilinear weights, hence the bits
    of ix, hence floor(ix) and the in-bounds flags, of torch's grid_sampler as the reference calls it."""
    ref = ns.multi_scale_deformable_attn_pytorch
    out = {"source": "utils.py:74-78 F.grid_sample(bilinear, zeros, align_corners=False) through "
                     "multi_scale_deformable_attn_pytorch on identity images; torch " + torch.__version__ + " CPU",
           "cases": {}}
    for W in (20, 40, 80, 160, 320, 13, 7):
        k = np.arange(-1, W + 2, dtype=np.float64)
        pts = []
        for base in ((k + 0.5) / W, k / W, (k + 0.25) / W):
            b32 = base.astype(np.float32)
            for d in (-2, -1, 0, 1, 2):
                v = b32.copy()
                for _ in range(abs(d)):
                    v = np.nextafter(v, np.float32(np.inf if d > 0 else -np.inf), dtype=np.float32)
                pts.append(v)
        rnd = (np.random.RandomState(W).rand(1500).astype(np.float32) * np.float32(1.2) - np.float32(0.1))
        pts = torch.from_numpy(np.concatenate(pts + [rnd]))
        n = pts.numel()
        # x-direction probe: level [1, W], y fixed at 0.5 (iy == 0 exactly -> weight 1 on row 0)
        loc = torch.zeros(1, n, 1, 1, 1, 2)
        loc[0, :, 0, 0, 0, 0] = pts
        loc[0, :, 0, 0, 0, 1] = 0.5
        ox = ref(torch.eye(W).view(1, W, 1, W), [[1, W]], loc, torch.ones(1, n, 1, 1, 1))[0]     # [n, W]
        # y-direction probe: level [W, 1]
        loc_y = torch.zeros(1, n, 1, 1, 1, 2)
        loc_y[0, :, 0, 0, 0, 1] = pts
        loc_y[0, :, 0, 0, 0, 0] = 0.5
        oy = ref(torch.eye(W).view(1, W, 1, W), [[W, 1]], loc_y, torch.ones(1, n, 1, 1, 1))[0]
        out["cases"][W] = dict(pts=pts, x_sparse=ox.to_sparse_coo(), y_sparse=oy.to_sparse_coo())
    _save("msda_index_probe", out)


RAGGED_CASES = {
    # name: (seed, B, Lq, H, Dh, shapes)
    "tiny_nonsquare": (41, 2, 19, 4, 8, [[9, 13], [5, 7], [3, 4]]),
    "small_dh32": (42, 2, 50, 8, 32, [[20, 20], [10, 10], [5, 5]]),
    "small_dh64": (43, 1, 40, 8, 64, [[24, 16], [12, 8], [6, 4]]),
}


def gen_ragged(ns):
    """multi_scale_deformable_attn_pytorch_cls / _box (utils.py:92-191): 2/4/6 and 6/4/2 points per level."""
    out = {"source": "ultralytics/nn/modules/utils.py:92-140 (_cls, points 2/4/6) and :143-191 (_box, points 6/4/2), torch "
                     + torch.__version__ + " CPU fp32", "cases": {}}
    for kind, points, ref in (("cls", (2, 4, 6), ns.utils.multi_scale_deformable_attn_pytorch_cls),
                              ("box", (6, 4, 2), ns.utils.multi_scale_deformable_attn_pytorch_box)):
        for name, (seed, B, Lq, H, Dh, shapes) in RAGGED_CASES.items():
            value, loc, attn, grad_out = msda.make_ragged_inputs(seed, B, Lq, H, Dh, shapes, points)
            value.requires_grad_(), loc.requires_grad_(), attn.requires_grad_()
            o = ref(value, shapes, loc, attn)
            o.backward(grad_out)
            case = dict(seed=seed, B=B, Lq=Lq, H=H, Dh=Dh, shapes=shapes, points=points, out=o.detach().clone(),
                        grad_loc=loc.grad.clone(), grad_attn=attn.grad.clone())
            if value.numel() <= 60_000:
                case["grad_value"] = value.grad.clone()
            else:
                case["grad_value_subset"] = _subset(value.grad, 8000, seed + 1000)
                case["grad_value_norm"] = value.grad.double().norm().item()
            out["cases"][kind + "_" + name] = case
    # the two attention modules and the decoupled decoder layer that calls them (transformer.py:300-495, 561-658)
    shapes = [[20, 20], [10, 10], [5, 5]]
    d, H, B, Lq = 256, 8, 2, 48
    torch.manual_seed(0)
    layer = ns.transformer.DecouplingDeformableTransformerDecoderLayer(d, H, 512, 0.0, torch.nn.ReLU(), 3, 4)
    manifest = seeding.seeded_fill(layer, 44)
    Lv = sum(h * w for h, w in shapes)
    embed = seeding.seeded_tensor(45, "embed", (B, Lq, d)).requires_grad_()
    embed1 = seeding.seeded_tensor(45, "embed1", (B, Lq, d)).requires_grad_()
    # smooth maps: with white-noise features a sample within rounding distance of a cell edge flips a whole gradient row
    feats = seeding.seeded_smooth_tokens(45, "feats", B, d, shapes, factor=2).requires_grad_()
    ref_box = torch.cat([seeding.seeded_uniform(45, "ref_xy", (B, Lq, 2)),
                         seeding.seeded_uniform(45, "ref_wh", (B, Lq, 2), 0.01, 0.3)], -1)
    pos = seeding.seeded_tensor(45, "pos", (B, Lq, d))
    mask = torch.zeros(Lq, Lq, dtype=torch.bool)
    mask[:16, 16:] = True
    mask[16:, :16] = True
    o_cls, o_box = layer(embed, embed1, ref_box, feats, shapes, None, mask, pos)
    (_probe_loss(o_cls, 46, "p_cls") + _probe_loss(o_box, 46, "p_box")).backward()
    out["layer"] = dict(manifest=manifest, fill_seed=44, d=d, H=H, B=B, Lq=Lq, d_ffn=512, shapes=shapes,
                        out_cls=o_cls.detach().clone(), out_box=o_box.detach().clone(),
                        grad_embed=embed.grad.clone(), grad_embed1=embed1.grad.clone(),
                        grad_feats_subset=_subset(feats.grad, 8000, 48), grad_feats_norm=feats.grad.double().norm().item(),
                        # small parameters in full, the big matrices as fixed random subsets
                        param_grads={k: (v.grad.clone() if v.numel() <= 4096 else _subset(v.grad, 2000, 47))
                                     for k, v in layer.named_parameters() if v.grad is not None})
    _save("msda_ragged", out)


# ------------------------------------------------------------------------------------------------ modules
def _probe_loss(out, seed, name):
    return (out * seeding.seeded_tensor(seed, name, out.shape)).sum()


def _msda_inputs(seed, B, Lq, d, shapes, ref_dim=4):
    Lv = sum(h * w for h, w in shapes)
    query = seeding.seeded_tensor(seed, "query", (B, Lq, d))
    value = seeding.seeded_tensor(seed, "value", (B, Lv, d))
    if ref_dim == 4:
        ref = torch.cat([seeding.seeded_uniform(seed, "ref_xy", (B, Lq, 1, 2)),
                         seeding.seeded_uniform(seed, "ref_wh", (B, Lq, 1, 2), 0.01, 0.3)], -1)
    else:
        ref = seeding.seeded_uniform(seed, "ref_xy", (B, Lq, len(shapes), 2))
    return query, ref, value


def gen_msdeform(ns):
    """MSDeformAttn (transformer.py:204-299): init-state KAT, seeded weights, 2-d reference points."""
    out = {"source": "ultralytics/nn/modules/transformer.py:204-299 MSDeformAttn", "cases": {}}
    shapes = [[20, 20], [10, 10], [5, 5]]
    for name, (d, H, fill, ref_dim) in {"init_state": (256, 8, None, 4), "seeded": (256, 8, 21, 4),
                                        "seeded_ref2": (256, 8, 22, 2), "seeded_d512": (512, 8, 23, 4)}.items():
        torch.manual_seed(0)
        m = ns.MSDeformAttn(d, 3, H, 4)
        manifest = seeding.seeded_fill(m, fill) if fill is not None else {k: tuple(v.shape) for k, v in m.state_dict().items()}
        if fill is None:   # KAT: only the Xavier projections are random -> fix them by seed as well
            with torch.no_grad():
                for k in ("value_proj", "output_proj"):
                    getattr(m, k).weight.copy_(seeding.seeded_tensor(20, k, getattr(m, k).weight.shape) / d ** 0.5)
        query, ref, value = _msda_inputs(30, 2, 50, d, shapes, ref_dim)
        query.requires_grad_(), value.requires_grad_()
        o = m(query, ref, value, shapes)
        _probe_loss(o, 31, "probe").backward()
        out["cases"][name] = dict(d=d, H=H, fill=fill, ref_dim=ref_dim, shapes=shapes, manifest=manifest,
                                  out=o.detach().clone(), grad_query=query.grad.clone(), grad_value_subset=_subset(value.grad, 8000, 33),
                                  grad_value_norm=value.grad.double().norm().item(),
                                  grad_param_norms={k: p.grad.double().norm().item() for k, p in m.named_parameters()},
                                  grad_param_subsets={k: _subset(p.grad, 2000, 32) for k, p in m.named_parameters()})
    _save("modules_msdeform", out)


def gen_layer(ns):
    """DeformableTransformerDecoderLayer (transformer.py:498-558)."""
    shapes = [[20, 20], [10, 10], [5, 5]]
    d, H, B, Lq = 256, 8, 2, 48
    torch.manual_seed(0)
    m = ns.DeformableTransformerDecoderLayer(d, H, 1024, 0., torch.nn.ReLU(), 3, 4)
    manifest = seeding.seeded_fill(m, 41)
    Lv = sum(h * w for h, w in shapes)
    embed = seeding.seeded_tensor(42, "embed", (B, Lq, d)).requires_grad_()
    feats = seeding.seeded_tensor(42, "feats", (B, Lv, d)).requires_grad_()
    pos = seeding.seeded_tensor(42, "pos", (B, Lq, d))
    ref = torch.cat([seeding.seeded_uniform(42, "xy", (B, Lq, 2)), seeding.seeded_uniform(42, "wh", (B, Lq, 2), 0.01, 0.3)], -1)
    mask = torch.zeros(Lq, Lq, dtype=torch.bool)
    mask[16:, :16] = True
    mask[:8, 8:16] = True
    mask[8:16, :8] = True
    o = m(embed, ref, feats, shapes, None, mask, pos)
    _probe_loss(o, 43, "probe").backward()
    _save("modules_layer", dict(source="ultralytics/nn/modules/transformer.py:498-558", d=d, H=H, B=B, Lq=Lq, shapes=shapes,
                                manifest=manifest, out=o.detach().clone(), grad_embed=embed.grad.clone(),
                                grad_feats=feats.grad.clone()))


def gen_contrastive(ns):
    """ContrastiveHeadMLP (block.py:522-541)."""
    out = {"source": "ultralytics/nn/modules/block.py:522-541 ContrastiveHeadMLP", "cases": {}}
    for K in (10, 80):
        m = ns.ContrastiveHeadMLP()
        x = seeding.seeded_tensor(50 + K, "x", (2, 300, 512)).requires_grad_()
        w = seeding.seeded_tensor(50 + K, "w", (2, K, 512)).requires_grad_()
        o = m(x, w)
        _probe_loss(o, 51, "probe").backward()
        out["cases"][K] = dict(out=o.detach().clone(), grad_x_subset=_subset(x.grad, 8000, 52),
                               grad_x_norm=x.grad.double().norm().item(), grad_w=w.grad.clone(),
                               grad_logit_scale=m.logit_scale.grad.clone(), grad_bias=m.bias.grad.clone())
    _save("modules_contrastive", out)


MAXSIG_CASES = {"c256_20x20_n10": (256, 8, 20, 20, 10, 2), "c128_24x16_n10": (128, 4, 24, 16, 10, 2),
                "c64_40x40_n10": (64, 2, 40, 40, 10, 1), "c256_40x40_n80": (256, 8, 40, 40, 80, 1)}


def gen_maxsigmoid(ns):
    """MaxSigmoidAttnBlock (extra_modules/block.py:194-226), as TIAGELAN builds it (c1 == c2 == ec, hc = 32)."""
    out = {"source": "ultralytics/nn/extra_modules/block.py:194-226 MaxSigmoidAttnBlock", "cases": {}}
    for name, (C, nh, Hh, Ww, N, B) in MAXSIG_CASES.items():
        m = ns.MaxSigmoidAttnBlock(C, C, nh=nh, ec=C)
        manifest = seeding.seeded_fill(m, 61)
        x0 = seeding.seeded_tensor(62, "x", (B, C, Hh, Ww))
        guide = seeding.seeded_tensor(62, "guide", (B, N, 512))
        case = dict(C=C, nh=nh, H=Hh, W=Ww, N=N, B=B, manifest=manifest)
        for mode in ("eval", "train"):
            m.train(mode == "train")
            m.zero_grad()
            x = x0.clone().requires_grad_()
            g = guide.clone().requires_grad_()
            o = m(x, g)
            _probe_loss(o, 63, "probe").backward()
            case[mode] = dict(out_subset=_subset(o, 6000, 64), out_norm=o.detach().double().norm().item(),
                              grad_x_subset=_subset(x.grad, 6000, 65), grad_x_norm=x.grad.double().norm().item(),
                              grad_guide=g.grad.clone(), grad_bias=m.bias.grad.clone(),
                              grad_gl_weight_norm=m.gl.weight.grad.double().norm().item())
        case["running_mean_after_train"] = m.proj_conv.bn.running_mean.clone()
        case["running_var_after_train"] = m.proj_conv.bn.running_var.clone()
        out["cases"][name] = case
    _save("modules_maxsigmoid", out)


def _synthetic_targets(seed, B, lo, hi, nc=10):
    g = torch.Generator().manual_seed(seed)
    groups = [int(torch.randint(lo, hi + 1, (1,), generator=g)) for _ in range(B)]
    n = sum(groups)
    boxes = torch.cat([torch.rand(n, 2, generator=g), 0.01 + 0.29 * torch.rand(n, 2, generator=g)], -1)
    cls = torch.randint(0, nc, (n,), generator=g)
    idx = torch.cat([torch.full((k,), i, dtype=torch.long) for i, k in enumerate(groups)])
    return {"cls": cls, "bboxes": boxes, "batch_idx": idx, "gt_groups": groups}


def gen_heads(ns):
    """RTDETRDecoder eval (BASELINE.json config 1 shapes) and ManbaWorldDecoder train+eval with a CDN group
    (TAMTR.yaml:67 shapes; VSSBlocks replaced by identity: their CUDA extension is not in the reference tree)."""
    import torch.nn as nn
    out = {"source": "ultralytics/nn/modules/head.py:174-435 RTDETRDecoder, :1005-1290 ManbaWorldDecoder "
                     "(VSSBlocks=identity), models/utils/ops.py:152-291 get_cdn_group", "cases": {}}
    # Multi-layer heads amplify fp32 rounding (the reference's own fp32 output is ~1e-3 away from its fp64 output
    # with these weights), so every head case stores the fp64 result as the target plus `ref32_err`, the distance of
    # the reference's fp32 run from it: the tests allow max(1e-4, 3 * ref32_err).
    def rel(a, b):
        return ((a.double() - b.double()).norm() / b.double().norm()).item()

    # --- RT-DETR head, eval, S-base
    def selection_margin(m, xs, nq):
        """gap between the last selected and the first rejected token score (head.py:1237): the goldens must not sit on
        a near-tie, or two correct implementations legitimately select different queries."""
        with torch.no_grad():
            feats, shapes = m._get_encoder_input(xs)
            anchors, valid = m._generate_anchors(shapes, dtype=feats.dtype, device=feats.device)
            sc = m.enc_score_head(m.enc_output(valid * feats)).max(-1).values.sort(1, descending=True).values
        return (sc[:, nq - 1] - sc[:, nq]).min().item()

    torch.manual_seed(0)
    m = ns.RTDETRDecoder(nc=10, ch=(256, 256, 256)).eval()
    manifest = seeding.seeded_fill(m, 71)
    for in_seed in range(72, 200):
        xs = [seeding.seeded_tensor(in_seed, f"x{i}", (2, 256, s, s)) for i, s in enumerate((80, 40, 20))]
        if selection_margin(m, xs, 300) > 2e-3:
            break
    rt_seed = in_seed
    with torch.no_grad():
        y32, (db32, ds32, eb32, es32, _) = m(xs)
        m.double()
        y, (db, ds, eb, es, _) = m([x.double() for x in xs])
    out["cases"]["rtdetr_eval_sbase"] = dict(
        manifest=manifest, input_seed=rt_seed, y=y.float(), dec_bboxes=db.float(), dec_scores=ds.float(), enc_bboxes=eb.float(),
        enc_scores=es.float(),
        ref32_err=dict(y=rel(y32, y), dec_bboxes=rel(db32, db), dec_scores=rel(ds32, ds), enc_bboxes=rel(eb32, eb),
                       enc_scores=rel(es32, es)))

    # --- MEH head (ManbaWorldDecoder), TAMTR.yaml:67 channel/width config
    class _IdentityVSS(nn.Identity):
        def __init__(self, *a, **k):
            super().__init__()
    orig = ns.head.VSSBlock
    ns.head.VSSBlock = _IdentityVSS
    from .head_ref import surrogate_loss
    orig_cdn = ns.ops.get_cdn_group

    def cdn_any_dtype(batch, nc, nq, class_embed, *a, **k):
        # the reference's get_cdn_group writes into fp32 `torch.zeros` buffers (ops.py:243-244) and so cannot run
        # under a float64 model; run it in fp32 (identical queries) and cast the result for the fp64 target run
        if batch is not None:
            batch = dict(batch, bboxes=batch["bboxes"].float())
        e, b, mask, meta = orig_cdn(batch, nc, nq, class_embed.float(), *a, **k)
        return (None if e is None else e.to(class_embed.dtype)), (None if b is None else b.to(class_embed.dtype)), mask, meta
    ns.ops.get_cdn_group = cdn_any_dtype
    try:
        # batch >= 2 on purpose: torch's CPU BatchNorm backward is wrong at batch 1 for this op sequence
        # (oracle/head_ref.py::_ContiguousGrad), so a batch-1 golden of the reference's CPU path would pin a bug.
        for name, sizes, B in (("meh_syaml_small", (40, 20, 10), 2), ("meh_syaml_full", (160, 80, 40), 2)):
            torch.manual_seed(0)
            m = ns.ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3)
            manifest = seeding.seeded_fill(m, 73)
            filled = {k: v.clone() for k, v in m.state_dict().items()}
            for in_seed in range(74, 200):
                xs = [seeding.seeded_smooth_map(in_seed, f"x{i}", (B, c, s, s)) for i, (c, s) in enumerate(zip((128, 256, 512), sizes))]
                m.train()
                ok_train = selection_margin(m, xs, 100) > 2e-3      # (batch statistics differ between train and eval)
                m.eval()
                if ok_train and selection_margin(m, xs, 100) > 2e-3:
                    break
            m.load_state_dict(filled)       # the train-mode margin probes above moved the BN running statistics
            m.train()
            text = F_normalize(seeding.seeded_tensor(74, "text", (B, 10, 512)))
            batch = _synthetic_targets(75, B, 5, 20)
            case = dict(manifest=manifest, sizes=sizes, B=B, batch=batch, input_seed=in_seed)
            runs = {}
            sd0 = {k: v.clone() for k, v in m.state_dict().items()}
            for prec in ("f32", "f64"):
                dt = torch.float32 if prec == "f32" else torch.float64
                m.float().load_state_dict(sd0)      # the train-mode forward below updates BN running stats
                m.to(dt).train()
                m.zero_grad()
                torch.manual_seed(1234)
                dn_embed, dn_bbox, attn_mask, dn_meta = ns.ops.get_cdn_group(batch, 10, 100, m.denoising_class_embed.weight,
                                                                         100, 0.5, 1.0, True)
                if prec == "f32":
                    case["cdn"] = dict(dn_bbox=dn_bbox.clone(), attn_mask=attn_mask.clone(), dn_meta=dn_meta,
                                       dn_embed_norm=dn_embed.norm().item())
                # the reference builds the denoising boxes in fp32 whatever the model dtype (ops.py:243-244
                # `torch.zeros(...)` without dtype) -> identical queries in both precisions
                torch.manual_seed(1234)
                xs_g = [x.detach().to(dt).clone().requires_grad_() for x in xs]
                b2 = dict(batch, bboxes=batch["bboxes"].to(dt)) if prec == "f64" else batch
                db, ds, eb, es, meta = m(xs_g, text.to(dt), b2)
                loss = surrogate_loss(db, ds, eb, es)
                loss.backward()
                m.eval()
                with torch.no_grad():
                    y, _ = m([x.to(dt) for x in xs], text.to(dt))
                runs[prec] = dict(dec_bboxes=db.detach(), dec_scores=ds.detach(), enc_bboxes=eb.detach(),
                                  enc_scores=es.detach(), loss=loss.item(), grad_x=[x.grad for x in xs_g], eval_y=y,
                                  grad_params={k: p.grad for k, p in m.named_parameters() if p.grad is not None})
            r32, r64 = runs["f32"], runs["f64"]
            case["train"] = dict(
                dec_bboxes=r64["dec_bboxes"].float(), dec_scores=r64["dec_scores"].float(),
                enc_bboxes=r64["enc_bboxes"].float(), enc_scores=r64["enc_scores"].float(), loss=r64["loss"],
                grad_x_norms=[g.norm().item() for g in r64["grad_x"]],
                grad_x2_subset=_subset(r64["grad_x"][2].float(), 6000, 76),
                grad_param_norms={k: g.norm().item() for k, g in r64["grad_params"].items()},
                ref32_err=dict(dec_bboxes=rel(r32["dec_bboxes"], r64["dec_bboxes"]),
                               dec_scores=rel(r32["dec_scores"], r64["dec_scores"]),
                               enc_bboxes=rel(r32["enc_bboxes"], r64["enc_bboxes"]),
                               enc_scores=rel(r32["enc_scores"], r64["enc_scores"]),
                               loss=abs(r32["loss"] - r64["loss"]) / abs(r64["loss"]),
                               grad_x2=rel(r32["grad_x"][2], r64["grad_x"][2]),
                               grad_params=max(rel(r32["grad_params"][k], g) for k, g in r64["grad_params"].items()
                                               if g.norm() > 0)))
            case["eval_y"] = r64["eval_y"].float()
            case["eval_ref32_err"] = rel(r32["eval_y"], r64["eval_y"])
            out["cases"][name] = case
    finally:
        ns.head.VSSBlock = orig
        ns.ops.get_cdn_group = orig_cdn
    _save("modules_heads", out)


def F_normalize(t):
    return torch.nn.functional.normalize(t, dim=-1)


GENERATORS = {"core": gen_core, "index_probe": gen_index_probe, "msdeform": gen_msdeform, "layer": gen_layer,
              "contrastive": gen_contrastive, "maxsigmoid": gen_maxsigmoid, "heads": gen_heads, "ragged": gen_ragged}


def main(argv):
    ns = reference_loader.hot_path()
    torch.manual_seed(0)
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    names = argv or list(GENERATORS)
    for n in names:
        GENERATORS[n](ns)


if __name__ == "__main__":
    main(sys.argv[1:])
