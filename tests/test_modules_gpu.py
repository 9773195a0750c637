"""GPU parity tests of the module-level drop-ins (tamtr_b200.modules / tamtr_b200.head) against goldens produced by
the reference's own modules and against the CPU oracle.  fp32 <= 1e-4 relative (multi-layer heads: see head_tol),
bf16 <= 2e-2 relative against the fp32 reference on the same weights."""
import pytest
import torch

from helpers import (align_queries, check_full_or_subset, filled_state_dict, gather_rows, load_golden, probe_loss,
                     rel_l2, subset_err)
from oracle import head_ref, seeding
from oracle.make_goldens import MAXSIG_CASES, _msda_inputs, _synthetic_targets

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2


def head_tol(ref32_err, floor=FP32_TOL, factor=3.0):
    """Multi-layer heads amplify fp32 rounding: the golden target is the reference's fp64 output and `ref32_err`
    the distance of the reference's own fp32 run from it.  Allowed: the north star's 1e-4, or `factor` x that."""
    return max(floor, factor * ref32_err)


def _cpu_sd(module, prefix=""):
    return {prefix + k: v.detach().float().cpu() for k, v in module.state_dict().items()}


# ---------------------------------------------------------------------------------------------- MSDeformAttn
@pytest.mark.parametrize("name", ["init_state", "seeded", "seeded_ref2", "seeded_d512"])
def test_msdeform_attn_fp32(cuda_lib, name):
    from tamtr_b200.modules import MSDeformAttn
    c = load_golden("modules_msdeform")["cases"][name]
    d = c["d"]
    m = MSDeformAttn(d, 3, c["H"], 4)
    if c["fill"] is None:
        assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == c["manifest"]
        with torch.no_grad():
            for k in ("value_proj", "output_proj"):
                getattr(m, k).weight.copy_(seeding.seeded_tensor(20, k, getattr(m, k).weight.shape) / d ** 0.5)
    else:
        filled_state_dict(m, c["fill"], c["manifest"])
    m.cuda()
    query, ref, value = _msda_inputs(30, 2, 50, d, c["shapes"], c["ref_dim"])
    query, value = query.cuda().requires_grad_(), value.cuda().requires_grad_()
    out = m(query, ref.cuda(), value, c["shapes"])
    probe_loss(out, 31, "probe").backward()
    assert rel_l2(out, c["out"]) < FP32_TOL
    assert rel_l2(query.grad, c["grad_query"]) < FP32_TOL
    check_full_or_subset(value.grad, c, "grad_value", FP32_TOL)
    for k, p in m.named_parameters():
        n = c["grad_param_norms"][k]
        assert abs(p.grad.double().norm().item() - n) <= FP32_TOL * n + 1e-9, k
        if n > 0:
            assert subset_err(p.grad, c["grad_param_subsets"][k]) < FP32_TOL, k


@pytest.mark.parametrize("mode", ["autocast", "module_bf16"])
def test_msdeform_attn_bf16(cuda_lib, mode):
    from tamtr_b200.modules import MSDeformAttn
    c = load_golden("modules_msdeform")["cases"]["seeded_d512"]
    m = MSDeformAttn(512, 3, 8, 4)
    filled_state_dict(m, c["fill"], c["manifest"])
    sd = _cpu_sd(m, "m.")
    query, ref, value = _msda_inputs(30, 2, 50, 512, c["shapes"], 4)
    # fp32 reference on bf16-rounded inputs and weights
    sd_r = {k: v.bfloat16().float().requires_grad_() for k, v in sd.items()}
    q_r, v_r = query.bfloat16().float().requires_grad_(), value.bfloat16().float().requires_grad_()
    out_r = head_ref.msdeform_attn(sd_r, "m", q_r, ref, v_r, c["shapes"], 8)
    probe_loss(out_r, 31, "probe").backward()
    m.cuda()
    if mode == "module_bf16":
        m.bfloat16()
        q, v = query.cuda().bfloat16().requires_grad_(), value.cuda().bfloat16().requires_grad_()
        out = m(q, ref.cuda(), v, c["shapes"])
    else:
        q, v = query.cuda().requires_grad_(), value.cuda().requires_grad_()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(q, ref.cuda(), v, c["shapes"])
    assert out.dtype == torch.bfloat16
    probe_loss(out.float(), 31, "probe").backward()
    assert rel_l2(out, out_r) < BF16_TOL
    assert rel_l2(q.grad, q_r.grad) < BF16_TOL
    assert rel_l2(v.grad, v_r.grad) < BF16_TOL
    assert rel_l2(m.value_proj.weight.grad, sd_r["m.value_proj.weight"].grad) < BF16_TOL
    # d/d(location) differentiates the bilinear interpolant, i.e. takes DIFFERENCES of neighbouring values; at module
    # level the projected `value` is itself rounded to bf16 (the fp32 reference keeps it fp32), which this gradient
    # amplifies.  The op-level contract (same rounded inputs on both sides, tests/test_msda_gpu.py) holds at 2e-2.
    err = rel_l2(m.sampling_offsets.weight.grad, sd_r["m.sampling_offsets.weight"].grad)
    print(f"bf16 {mode}: sampling_offsets.weight.grad rel-L2 = {err:.4f}")
    assert err < 2.5 * BF16_TOL


def test_locations_and_weights_match_reference_arithmetic(cuda_lib):
    """The fused projection epilogue must hand the sampler (almost) the same locations the reference computes:
    same op order and per-op rounding, only the GEMM accumulation order differs (~1e-7)."""
    from tamtr_b200 import ops
    c = load_golden("modules_msdeform")["cases"]["seeded"]
    from tamtr_b200.modules import MSDeformAttn
    m = MSDeformAttn(256, 3, 8, 4)
    filled_state_dict(m, c["fill"], c["manifest"])
    sd = _cpu_sd(m, "m.")
    query, ref, value = _msda_inputs(30, 2, 50, 256, c["shapes"], 4)
    _, loc_r, aw_r = head_ref.msdeform_attn(sd, "m", query, ref, value, c["shapes"], 8, return_aux=True)
    m.cuda()
    loc, aw = ops.sampling_locations_and_weights(query.cuda(), ref.cuda(), m.sampling_offsets.weight,
                                                 m.sampling_offsets.bias, m.attention_weights.weight,
                                                 m.attention_weights.bias, c["shapes"], 8, 3, 4)
    assert (loc.cpu() - loc_r).abs().max() < 2e-6 and (aw.cpu() - aw_r).abs().max() < 1e-6
    assert torch.allclose(aw.sum((-1, -2)).cpu(), torch.ones(2, 50, 8), atol=1e-6)
    with pytest.raises((ValueError, RuntimeError)):       # transformer.py:295
        ops.sampling_locations_and_weights(query.cuda(), torch.rand(2, 50, 1, 3).cuda(), m.sampling_offsets.weight,
                                           m.sampling_offsets.bias, m.attention_weights.weight,
                                           m.attention_weights.bias, c["shapes"], 8, 3, 4)


# ---------------------------------------------------------------------------------------------- decoder layer
def test_decoder_layer_fp32(cuda_lib):
    from tamtr_b200.modules import DeformableTransformerDecoderLayer
    c = load_golden("modules_layer")
    m = DeformableTransformerDecoderLayer(c["d"], c["H"], 1024, 0., torch.nn.ReLU(), 3, 4)
    filled_state_dict(m, 41, c["manifest"])
    m.cuda()
    B, Lq, d = c["B"], c["Lq"], c["d"]
    Lv = sum(h * w for h, w in c["shapes"])
    embed = seeding.seeded_tensor(42, "embed", (B, Lq, d)).cuda().requires_grad_()
    feats = seeding.seeded_tensor(42, "feats", (B, Lv, d)).cuda().requires_grad_()
    pos = seeding.seeded_tensor(42, "pos", (B, Lq, d)).cuda()
    ref = torch.cat([seeding.seeded_uniform(42, "xy", (B, Lq, 2)), seeding.seeded_uniform(42, "wh", (B, Lq, 2), 0.01, 0.3)], -1).cuda()
    mask = torch.zeros(Lq, Lq, dtype=torch.bool)
    mask[16:, :16] = True
    mask[:8, 8:16] = True
    mask[8:16, :8] = True
    out = m(embed, ref, feats, c["shapes"], None, mask.cuda(), pos)
    probe_loss(out, 43, "probe").backward()
    assert rel_l2(out, c["out"]) < FP32_TOL
    assert rel_l2(embed.grad, c["grad_embed"]) < FP32_TOL and rel_l2(feats.grad, c["grad_feats"]) < FP32_TOL


# ---------------------------------------------------------------------------------------------- contrastive head
@pytest.mark.parametrize("K", [10, 80])
def test_contrastive_head(cuda_lib, K):
    from tamtr_b200.modules import ContrastiveHeadMLP
    c = load_golden("modules_contrastive")["cases"][K]
    m = ContrastiveHeadMLP().cuda()
    x = seeding.seeded_tensor(50 + K, "x", (2, 300, 512)).cuda().requires_grad_()
    w = seeding.seeded_tensor(50 + K, "w", (2, K, 512)).cuda().requires_grad_()
    out = m(x, w)
    probe_loss(out, 51, "probe").backward()
    assert rel_l2(out, c["out"]) < FP32_TOL
    check_full_or_subset(x.grad, c, "grad_x", FP32_TOL)
    assert rel_l2(w.grad, c["grad_w"]) < FP32_TOL
    assert rel_l2(m.logit_scale.grad, c["grad_logit_scale"]) < FP32_TOL
    assert rel_l2(m.bias.grad, c["grad_bias"]) < FP32_TOL
    # bf16 activations
    xb = x.detach().bfloat16().requires_grad_()
    ob = m(xb, w.detach())
    ref = head_ref.contrastive_head({"h.logit_scale": m.logit_scale.detach().cpu(), "h.bias": m.bias.detach().cpu()}, "h",
                                    xb.detach().float().cpu(), w.detach().cpu())
    assert ob.dtype == torch.bfloat16 and rel_l2(ob, ref) < BF16_TOL


# ---------------------------------------------------------------------------------------------- max-sigmoid attention
@pytest.mark.parametrize("name", list(MAXSIG_CASES))
def test_max_sigmoid_attn_block(cuda_lib, name):
    from tamtr_b200.modules import MaxSigmoidAttnBlock
    c = load_golden("modules_maxsigmoid")["cases"][name]
    C, nh, Hh, Ww, N, B = MAXSIG_CASES[name]
    m = MaxSigmoidAttnBlock(C, C, nh=nh, ec=C)
    filled_state_dict(m, 61, c["manifest"])
    m.cuda()
    x0 = seeding.seeded_tensor(62, "x", (B, C, Hh, Ww)).cuda()
    guide = seeding.seeded_tensor(62, "guide", (B, N, 512)).cuda()
    for mode in ("eval", "train"):
        m.train(mode == "train")
        m.zero_grad()
        x = x0.clone().requires_grad_()
        g = guide.clone().requires_grad_()
        out = m(x, g)
        probe_loss(out, 63, "probe").backward()
        gold = c[mode]
        assert subset_err(out, gold["out_subset"]) < FP32_TOL
        assert abs(out.double().norm().item() - gold["out_norm"]) < FP32_TOL * gold["out_norm"]
        assert subset_err(x.grad, gold["grad_x_subset"]) < 5e-4      # conv+BN backward in cuDNN (TF32 off) vs CPU
        assert rel_l2(g.grad, gold["grad_guide"]) < 5e-4
        assert rel_l2(m.bias.grad, gold["grad_bias"]) < 5e-4
    assert rel_l2(m.proj_conv.bn.running_mean, c["running_mean_after_train"]) < FP32_TOL
    assert rel_l2(m.proj_conv.bn.running_var, c["running_var_after_train"]) < FP32_TOL


def test_max_sigmoid_gate_vs_oracle_bf16(cuda_lib):
    from tamtr_b200 import ops
    B, nh, hc, Hh, Ww, N = 2, 8, 32, 40, 40, 20
    x = seeding.seeded_tensor(1, "x", (B, nh * hc, Hh, Ww))
    g = seeding.seeded_tensor(1, "g", (B, N, nh, hc)) * 0.3
    bias = seeding.seeded_tensor(1, "b", (nh,))
    xb = x.bfloat16()
    aw = ops.max_sigmoid_gate(xb.cuda(), g.cuda(), bias.cuda(), nh, use_tensor_cores=False)   # CUDA-core path
    e = xb.float().view(B, nh, hc, Hh, Ww)
    ref = (torch.einsum("bmchw,bnmc->bmhwn", e, g).max(-1)[0] / hc ** 0.5 + bias[None, :, None, None]).sigmoid()
    assert rel_l2(aw, ref) < 1e-5        # bf16 storage, fp32 arithmetic: exact up to accumulation order


# ---------------------------------------------------------------------------------------------- heads
def test_rtdetr_head_eval_sbase(cuda_lib):
    """BASELINE.json config 1: RTDETRDecoder forward, batch 2, 80^2/40^2/20^2, d=256, 8 heads, 300 queries."""
    from tamtr_b200.head import RTDETRDecoder
    c = load_golden("modules_heads")["cases"]["rtdetr_eval_sbase"]
    m = RTDETRDecoder(nc=10, ch=(256, 256, 256)).eval()
    filled_state_dict(m, 71, c["manifest"])
    m.cuda()
    xs = [seeding.seeded_tensor(c["input_seed"], f"x{i}", (2, 256, s, s)).cuda() for i, s in enumerate((80, 40, 20))]
    with torch.no_grad():
        y, (db, ds, eb, es, _) = m(xs)
    e = c["ref32_err"]
    assert y.shape == (2, 300, 14)
    idx, ok = align_queries(eb, es, c["enc_bboxes"], c["enc_scores"])
    assert ok.float().mean() > 0.99, "query selection differs from the reference by more than boundary ties"
    gather = lambda t: gather_rows(t, idx)   # noqa: E731
    sel = ok.unsqueeze(-1)
    slack = 1.0 if bool(ok.all()) else 3.0      # a swapped query perturbs the others through self-attention
    assert rel_l2(eb.cpu() * sel, gather(c["enc_bboxes"]) * sel) < FP32_TOL
    assert rel_l2(es.cpu() * sel, gather(c["enc_scores"]) * sel) < FP32_TOL
    assert rel_l2(db[0].cpu() * sel, gather(c["dec_bboxes"][0]) * sel) < slack * head_tol(e["dec_bboxes"])
    assert rel_l2(ds[0].cpu() * sel, gather(c["dec_scores"][0]) * sel) < slack * head_tol(e["dec_scores"])
    assert rel_l2(y.cpu() * sel, gather(c["y"]) * sel) < slack * head_tol(e["y"])


@pytest.mark.parametrize("name", ["meh_syaml_small", "meh_syaml_full"])
def test_meh_head_train_and_eval(cuda_lib, name):
    """TAM-TR's head (ManbaWorldDecoder, TAMTR.yaml:67 config, VSSBlocks = identity on both sides), train mode with a
    denoising group, forward + backward; then eval."""
    from tamtr_b200.head import ManbaWorldDecoder
    c = load_golden("modules_heads")["cases"][name]
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False)
    filled_state_dict(m, 73, c["manifest"])
    m.cuda().train()
    B, sizes = c["B"], c["sizes"]
    xs = [seeding.seeded_smooth_map(c["input_seed"], f"x{i}", (B, ch, s, s)).cuda().requires_grad_()
          for i, (ch, s) in enumerate(zip((128, 256, 512), sizes))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(74, "text", (B, 10, 512)), dim=-1).cuda()
    batch = _synthetic_targets(75, B, 5, 20)        # CPU targets -> the CDN RNG stream matches the golden's
    torch.manual_seed(1234)
    db, ds, eb, es, meta = m(xs, text, batch)
    g, e = c["train"], c["train"]["ref32_err"]
    assert meta["dn_num_split"] == c["cdn"]["dn_meta"]["dn_num_split"]
    idx, ok = align_queries(eb, es, g["enc_bboxes"], g["enc_scores"])     # top-k rows may swap on near ties
    assert bool(ok.all())
    n_dn = meta["dn_num_split"][0]

    def aligned(t):          # golden [layers, B, n_dn + nq, C] with its selected rows permuted into our order
        tail = torch.stack([gather_rows(t[i][:, n_dn:], idx) for i in range(t.shape[0])])
        return torch.cat([t[:, :, :n_dn], tail], 2)
    assert rel_l2(eb, gather_rows(g["enc_bboxes"], idx)) < FP32_TOL and rel_l2(es, gather_rows(g["enc_scores"], idx)) < FP32_TOL
    assert rel_l2(db, aligned(g["dec_bboxes"])) < head_tol(e["dec_bboxes"])
    assert rel_l2(ds, aligned(g["dec_scores"])) < head_tol(e["dec_scores"])
    loss = head_ref.surrogate_loss(db, ds, eb, es)
    assert abs(loss.item() - g["loss"]) < 1e-4 * abs(g["loss"])
    loss.backward()
    # Gradients w.r.t. the raw feature maps pass through every fp32 library kernel of the head (GEMMs, SDPA, norm
    # reductions over up to 51 200 elements): on the GPU even the reference's own op sequence (oracle restatement run
    # with plain torch CUDA ops) is further from the fp64 target than the CPU run was.  Measure that distance here
    # and allow 3x of it (never less than the north star's 1e-4).
    sd_gpu = {k: (v.detach().clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.detach().clone())
              for k, v in m.state_dict().items()}
    xs_t = [x.detach().clone().requires_grad_() for x in xs]
    torch.manual_seed(1234)
    from tamtr_b200.head import get_cdn_group
    cdn = get_cdn_group(batch, 10, 100, sd_gpu["denoising_class_embed.weight"], 100, 0.5, 1.0, True)[:3]
    out_t = head_ref.head(sd_gpu, "", xs_t, 100, 3, 8, training=True, text=text, cdn=cdn)
    head_ref.surrogate_loss(*out_t).backward()
    err_torch = subset_err(xs_t[2].grad, g["grad_x2_subset"])
    err_ours = subset_err(xs[2].grad, g["grad_x2_subset"])
    print(f"{name}: grad_x2 subset rel-L2 vs fp64 reference: ours {err_ours:.2e}, torch CUDA ops {err_torch:.2e}, "
          f"reference CPU fp32 {e['grad_x2']:.2e}")
    gtol = max(FP32_TOL, 3.0 * err_torch, 6.0 * e["grad_x2"])
    for x, n in zip(xs, g["grad_x_norms"]):
        assert abs(x.grad.double().norm().item() - n) < gtol * n
    assert err_ours < gtol
    for k, p in m.named_parameters():
        n = g["grad_param_norms"].get(k, 0.0)
        if n > 0:
            assert abs(p.grad.double().norm().item() - n) < max(gtol, head_tol(e["grad_params"], factor=6.0)) * n, k
    m.eval()
    with torch.no_grad():
        y, (_, _, eb2, es2, _) = m([x.detach() for x in xs], text)
    # eval uses running statistics -> a different selection than in train mode: align through the boxes in y itself
    dist, idx2 = torch.cdist(y[..., :4].double().cpu(), c["eval_y"][..., :4].double()).min(-1)
    assert bool((dist < 1e-3).all())
    assert rel_l2(y, gather_rows(c["eval_y"], idx2)) < head_tol(c["eval_ref32_err"])


def test_text_decoder_bf16_autocast(cuda_lib):
    """bf16 path of the 3-layer text decoder (autocast: GEMMs and the sampler's value/out in bf16; index math, softmax
    and norms fp32).  The decoder is driven directly so that the comparison is not scrambled by a different top-k
    query selection under bf16 scores.  Outputs: the north star's 2e-2 against the fp32 run (which other tests pin to
    the reference).  Gradients through three stacked layers are chaotic under bf16 with these random weights and
    white-noise features -- the reference's OWN op sequence (oracle restatement run on the GPU under the same
    autocast: plain torch ops, fp32 grid_sample) deviates from its fp32 run by 18 % here -- so the criterion for
    gradients is: not worse than that (x1.25)."""
    from tamtr_b200.head import ManbaWorldDecoder
    c = load_golden("modules_heads")["cases"]["meh_syaml_small"]
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False)
    filled_state_dict(m, 73, c["manifest"])
    m.cuda().train()
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    B, Lq, shapes = 4, 120, [[40, 40], [20, 20], [10, 10]]
    Lv = sum(h * w for h, w in shapes)
    embed = seeding.seeded_tensor(81, "embed", (B, Lq, 512)).cuda()
    feats = seeding.seeded_tensor(81, "feats", (B, Lv, 512)).cuda()
    refer = torch.logit(torch.cat([seeding.seeded_uniform(81, "xy", (B, Lq, 2), 0.05, 0.95),
                                   seeding.seeded_uniform(81, "wh", (B, Lq, 2), 0.02, 0.3)], -1)).cuda()
    text = torch.nn.functional.normalize(seeding.seeded_tensor(81, "text", (B, 10, 512)), dim=-1).cuda()
    res = {}
    for impl in ("product", "torch_ops"):
        for mode in ("fp32", "bf16"):
            e, f = embed.clone().requires_grad_(), feats.clone().requires_grad_()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                if impl == "product":
                    db, ds = m.decoder(e, refer, f, shapes, text, m.dec_bbox_head, m.dec_score_head, m.query_pos_head)
                else:
                    db, ds = head_ref.decoder(sd, "", e, refer, f, shapes, 3, 8, True, text=text)
            (probe_loss(db.float(), 82, "pb") + 0.01 * probe_loss(ds.float(), 82, "ps")).backward()
            res[impl, mode] = (db.detach().float(), ds.detach().float(), e.grad.float(), f.grad.float())
    ours = [rel_l2(a, b) for a, b in zip(res["product", "bf16"], res["product", "fp32"])]
    theirs = [rel_l2(a, b) for a, b in zip(res["torch_ops", "bf16"], res["torch_ops", "fp32"])]
    print("bf16 vs fp32 rel-L2 (dec_bboxes, dec_scores, grad_embed, grad_feats): product", ours, "torch ops", theirs)
    assert rel_l2(res["product", "fp32"][0], res["torch_ops", "fp32"][0]) < FP32_TOL
    assert ours[0] < BF16_TOL and ours[1] < BF16_TOL
    assert ours[2] < max(BF16_TOL, 1.25 * theirs[2]) and ours[3] < max(BF16_TOL, 1.25 * theirs[3])


def test_modules_survive_deepcopy_pickle_and_half(cuda_lib):
    """SURVEY.md section 5 'Checkpoint': EMA deep-copies, checkpoints pickle .half() modules, validators call .float()."""
    import copy
    import io
    from tamtr_b200.modules import MSDeformAttn
    m = MSDeformAttn(256, 3, 8, 4).cuda()
    m2 = copy.deepcopy(m).half()
    buf = io.BytesIO()
    torch.save(m2, buf)
    buf.seek(0)
    m3 = torch.load(buf, weights_only=False).float()
    query, ref, value = _msda_inputs(30, 1, 10, 256, [[8, 8], [4, 4], [2, 2]], 4)
    a = m(query.cuda(), ref.cuda(), value.cuda(), [[8, 8], [4, 4], [2, 2]])
    b = m3(query.cuda(), ref.cuda(), value.cuda(), [[8, 8], [4, 4], [2, 2]])
    assert rel_l2(b, a) < 5e-3          # weights went through fp16
    h = m2(query.cuda().half(), ref.cuda().half(), value.cuda().half(), [[8, 8], [4, 4], [2, 2]])
    assert h.dtype == torch.float16 and rel_l2(h, a) < 2e-2


def test_meh_head_inference_1280(cuda_lib):
    """BASELINE.json config 5: high-resolution inference, 1280x1280 -> pyramid 320^2/160^2/80^2 (134 400 tokens),
    900 queries, one image per GPU.  The product head on the GPU against the oracle restatement of the reference's op
    sequence on the host CPU (fp32); rows are aligned through the predicted boxes (top-k may swap near-tied tokens)."""
    from tamtr_b200.head import ManbaWorldDecoder
    nq = 900
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, nq, 4, 8, 3, vss=False)
    sd = filled_state_dict(m, 81)
    m.cuda().eval()
    xs = [seeding.seeded_smooth_map(82, f"x{i}", (1, ch, s, s)) for i, (ch, s) in enumerate(zip((128, 256, 512), (320, 160, 80)))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(83, "text", (1, 10, 512)), dim=-1)
    with torch.no_grad():
        y, (db, ds, eb, es, _) = m([x.cuda() for x in xs], text.cuda())
        rb, rs, reb, res = head_ref.head({k: v.cpu() for k, v in sd.items()}, "", xs, nq, 3, 8, training=False, text=text)
    assert y.shape == (1, nq, 14) and db.shape[2] == nq
    y_ref = torch.cat([rb[-1], rs[-1].sigmoid()], -1)
    dist, idx = torch.cdist(y[..., :4].double().cpu(), y_ref[..., :4].double()).min(-1)
    ok = dist < 1e-3
    assert ok.float().mean() > 0.97, ok.float().mean()
    sel = ok.unsqueeze(-1)
    assert rel_l2(y.cpu() * sel, gather_rows(y_ref, idx) * sel) < 2e-2
