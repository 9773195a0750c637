"""GPU: the device-side contrastive denoising group (csrc/detloss.cu `tamtr_cdn_group`, head.cdn_group_device) and the
fixed-shape training step built on it (loss.DeviceTargets, dp.HeadTrainStep / dp.StepCache).

Pin: ultralytics/models/utils/ops.py:152-291.  The reference draws its random numbers with data-dependent shapes from
torch's global generator; the kernel takes ten uniforms per slot instead.  The first test replays the reference's draws of
the golden fixture (seed 1234: torch.rand / randint_like / randint_like / rand_like, in that order), hands the SAME
numbers to the kernel slot by slot, and must get the reference's own queries (tests/golden/modules_heads.pt, produced by
the unmodified reference) -- labels and attention mask exactly, boxes to the last bit before the logit."""
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import seeding
from oracle.make_goldens import _synthetic_targets

pytestmark = pytest.mark.gpu


def _replay_reference_draws(batch, nc, num_dn, capacity, seed, cls_noise_ratio=0.5):
    """The reference's RNG calls (ops.py:217-229) under `seed`, scattered into the kernel's [B, capacity, 10] layout.
    Returns (uniforms, expected labels [B, n_dn] with -1 in empty slots)."""
    groups = batch["gt_groups"]
    total, biggest = sum(groups), max(groups)
    ng = max(1, num_dn // biggest)
    B = len(groups)
    cls = batch["cls"].repeat(2 * ng)
    box = batch["bboxes"].repeat(2 * ng, 1)
    img = batch["batch_idx"].repeat(2 * ng).view(-1).long()
    slot = torch.cat([torch.arange(n, dtype=torch.long) for n in groups])
    slot = torch.cat([slot + biggest * i for i in range(2 * ng)])
    torch.manual_seed(seed)
    u_flip = torch.rand(cls.shape)
    where = torch.nonzero(u_flip < cls_noise_ratio * 0.5).squeeze(-1)
    new_label = torch.randint_like(where, 0, nc, dtype=cls.dtype)
    sign = torch.randint_like(box, 0, 2)
    part = torch.rand_like(box)
    uni = torch.full((B, capacity, 10), 0.5)
    uni[img, slot, 0] = u_flip
    lab = torch.zeros(cls.shape)
    lab[where] = (new_label.float() + 0.5) / nc             # the kernel's label = floor(u * nc)
    uni[img, slot, 1] = lab
    uni[img, slot, 2:6] = 0.25 + 0.5 * sign                 # < 0.5 -> -1, >= 0.5 -> +1
    uni[img, slot, 6:10] = part
    cls = cls.clone()
    cls[where] = new_label
    want = torch.full((B, 2 * biggest * ng), -1, dtype=torch.long)
    want[img, slot] = cls.long()
    return uni, want


def _run_kernel(batch, uni, nc, nq, num_dn, capacity, max_gt, ratio=0.5, scale=1.0):
    from tamtr_b200 import _lib
    from tamtr_b200.loss import DeviceTargets
    tgt = DeviceTargets(len(batch["gt_groups"]), max_gt, "cuda", capacity, num_dn).load(batch)
    B, D = tgt.bs, capacity
    dn_cls = torch.empty(B, D, dtype=torch.int64, device="cuda")
    dn_box = torch.empty(B, D, 4, device="cuda")
    valid = torch.empty(B, D, device="cuda")
    mask = torch.empty(D + nq, D + nq, dtype=torch.uint8, device="cuda")
    u = uni.cuda().contiguous()
    rc = _lib.lib().tamtr_cdn_group(tgt.boxes.data_ptr(), tgt.cls.data_ptr(), tgt.count.data_ptr(), u.data_ptr(),
                                    dn_cls.data_ptr(), dn_box.data_ptr(), valid.data_ptr(), mask.data_ptr(), B, max_gt, D, nq,
                                    nc, num_dn, ratio, scale, _lib.stream_ptr(torch.device("cuda")))
    _lib.check(rc, "cdn_group")
    torch.cuda.synchronize()
    return dn_cls.cpu(), dn_box.cpu(), valid.cpu(), mask.cpu().bool()


@pytest.mark.parametrize("capacity", [182, 200, 320])
def test_cdn_kernel_reproduces_the_reference_queries(cuda_lib, capacity):
    c = load_golden("modules_heads")["cases"]["meh_syaml_small"]
    B = c["B"]
    batch = _synthetic_targets(75, B, 5, 20)
    assert batch["gt_groups"] == c["batch"]["gt_groups"]
    gold_box, gold_mask = c["cdn"]["dn_bbox"], c["cdn"]["attn_mask"]
    n_dn, nq = c["cdn"]["dn_meta"]["dn_num_split"]
    assert capacity >= n_dn                                 # 182 = the exact fit: no padding slots at all
    uni, want_cls = _replay_reference_draws(batch, 10, 100, capacity, 1234)
    dn_cls, dn_box, valid, mask = _run_kernel(batch, uni, 10, nq, 100, capacity, max(batch["gt_groups"]) + 3)
    filled = want_cls >= 0
    # labels, occupancy and the mask: exact
    assert torch.equal(valid[:, :n_dn] > 0, filled) and not bool(valid[:, n_dn:].any())
    assert torch.equal(dn_cls[:, :n_dn][filled], want_cls[filled])
    assert not bool(dn_cls[:, :n_dn][~filled].any()) and not bool(dn_cls[:, n_dn:].any())
    real = torch.cat([torch.arange(n_dn), torch.arange(capacity, capacity + nq)])
    assert torch.equal(mask[real][:, real], gold_mask)
    # padding slots: invisible to every real query, see only themselves
    pad = torch.arange(n_dn, capacity)
    if len(pad):
        assert bool(mask[real][:, pad].all())
        own = mask[pad][:, pad]
        assert torch.equal(own, ~torch.eye(len(pad), dtype=torch.bool))
        assert bool(mask[pad][:, real].all())
    # boxes: the reference's values (the arithmetic before the logit is reproduced op by op; logf may differ by an ulp)
    assert not bool(dn_box[:, n_dn:].any()) and not bool(dn_box[:, :n_dn][~filled].any())
    got, ref = dn_box[:, :n_dn][filled], gold_box[filled]
    assert torch.allclose(got, ref, rtol=2e-6, atol=2e-6), (got - ref).abs().max()
    assert torch.equal(torch.sigmoid(got).round(decimals=5), torch.sigmoid(ref).round(decimals=5))


def test_cdn_kernel_without_noise_and_without_ground_truth(cuda_lib):
    """ops.py:217, 222: noise switched off leaves labels and (un-logit-ed!) boxes as they are; an empty batch yields no
    denoising query at all (ops.py:191-192 returns None: here every slot is padding)."""
    batch = _synthetic_targets(5, 3, 2, 9)
    cap, nq = 64, 30
    uni = torch.rand(3, cap, 10)
    dn_cls, dn_box, valid, mask = _run_kernel(batch, uni, 10, nq, 20, cap, 12, ratio=0.0, scale=0.0)
    groups, m = batch["gt_groups"], max(batch["gt_groups"])
    ng = max(1, 20 // m)
    start = 0
    for b, n in enumerate(groups):
        for copy in range(2 * ng):
            assert torch.equal(dn_box[b, copy * m:copy * m + n], batch["bboxes"][start:start + n])
            assert torch.equal(dn_cls[b, copy * m:copy * m + n], batch["cls"][start:start + n].long())
        start += n
    empty = {"cls": torch.zeros(0, dtype=torch.long), "bboxes": torch.zeros(0, 4), "batch_idx": torch.zeros(0, dtype=torch.long),
             "gt_groups": [0, 0, 0]}
    dn_cls, dn_box, valid, mask = _run_kernel(empty, uni, 10, nq, 20, cap, 12)
    assert not bool(valid.any()) and not bool(dn_box.any())
    assert bool(mask[cap:, :cap].all()) and not bool(mask[cap:, cap:].any())


def _head_and_inputs(B, sizes=(40, 20, 10), nd=100):
    from tamtr_b200.head import ManbaWorldDecoder
    torch.manual_seed(0)
    # noise off: the denoising queries are a deterministic function of the ground truth, so the host-planned and the
    # device-built group -- and an eager run and a graph replay -- must agree number for number
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False, nd=nd, label_noise_ratio=0.0,
                          box_noise_scale=0.0).cuda().train()
    xs = [seeding.seeded_smooth_map(5, f"x{i}", (B, c, s, s)).cuda() for i, (c, s) in enumerate(zip((128, 256, 512), sizes))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(5, "t", (B, 10, 512)), dim=-1).cuda()
    return m, xs, text


def _detection_loss(nc=10):
    from tamtr_b200.loss import RTDETRDetectionLoss
    crit = RTDETRDetectionLoss(nc=nc, use_vfl=True)

    def fn(out, targets):
        db, ds, eb, es, meta = out
        dnb, db = torch.split(db, meta["dn_num_split"], dim=2)
        dns, ds = torch.split(ds, meta["dn_num_split"], dim=2)
        db, ds = torch.cat([eb.unsqueeze(0), db]), torch.cat([es.unsqueeze(0), ds])
        return crit((db, ds), targets, dn_bboxes=dnb, dn_scores=dns, dn_meta=meta)
    return fn


def test_padded_device_group_equals_the_host_planned_group(cuda_lib):
    """fp32, noise off: the head + detection loss on a DeviceTargets bucket (padding slots, kernel-built group, targets
    derived from the counts inside the loss kernel) give the losses and parameter gradients of the reference-layout path
    (host-built group of the exact size, dn_meta with dn_pos_idx / dn_num_group)."""
    from tamtr_b200.loss import DeviceTargets
    B = 2
    m, xs, text = _head_and_inputs(B, nd=20)
    batch = _synthetic_targets(11, B, 3, 8)
    dev_batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    loss_fn = _detection_loss()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    results = []
    for kind in ("host", "device"):
        m.load_state_dict(sd)
        m.zero_grad()
        if kind == "host":
            out = m(xs, text, dev_batch)
            losses = loss_fn(out, dev_batch)
        else:
            tgt = DeviceTargets(B, 12, "cuda", 48, 20).load(batch)      # 48 slots for a 2 * 8 * 2 = 32-query group
            out = m(xs, text, tgt)
            assert out[0].shape[2] == 48 + 100
            losses = loss_fn(out, tgt)
        sum(losses.values()).backward()
        grads = torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel(), device="cuda") for p in m.parameters()])
        results.append(({k: float(v) for k, v in losses.items()}, grads))
    (l_host, g_host), (l_dev, g_dev) = results
    assert set(l_host) == set(l_dev)
    for k in l_host:
        assert abs(l_host[k] - l_dev[k]) <= 1e-4 * max(1.0, abs(l_host[k])), (k, l_host[k], l_dev[k])
    assert rel_l2(g_dev, g_host) < 1e-3


def test_one_captured_step_serves_batches_with_different_counts(cuda_lib):
    """dp.HeadTrainStep on a DeviceTargets: the graph captured on one batch, replayed after load_inputs() of another batch
    with other ground-truth counts, gives the eager result for THAT batch (loss and gradients); a batch that does not fit the
    bucket is refused; dp.StepCache captures a second bucket for it and keeps sharing the gradient buffer."""
    from tamtr_b200 import dp
    from tamtr_b200.loss import DeviceTargets
    B = 2
    m, xs, text = _head_and_inputs(B, nd=20)
    loss_fn = _detection_loss()
    scalar = (lambda out, static: sum(loss_fn(out, static[-1]).values()))
    batches = [_synthetic_targets(21 + i, B, lo, hi) for i, (lo, hi) in enumerate([(3, 8), (1, 4), (9, 12)])]
    sd = {k: v.clone() for k, v in m.state_dict().items()}

    def eager(batch, cap, slots):
        m.load_state_dict(sd)
        st = dp.HeadTrainStep(m, scalar, (xs, text, DeviceTargets(B, slots, "cuda", cap, 20).load(batch)), use_graph=False)
        loss = float(st.run())
        return loss, st.flat.flat.clone()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        want = [eager(b, 48, 12) for b in batches]
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()

    m.load_state_dict(sd)
    step = dp.HeadTrainStep(m, scalar, (xs, text, DeviceTargets(B, 12, "cuda", 48, 20).load(batches[0])), use_graph=True,
                            warmup=1)
    assert step.graph is not None
    for i in (0, 1, 2, 0):
        m.load_state_dict(sd)                       # (BatchNorm running statistics move with every step)
        step.load_inputs((None, None, batches[i]))
        loss = float(step.run())
        torch.cuda.synchronize()
        if not step.pack_grads:
            step.flat.gather()
        assert abs(loss - want[i][0]) <= 2e-4 * abs(want[i][0]), (i, loss, want[i][0])
        assert rel_l2(step.flat.flat, want[i][1]) < 2e-3, i
    with pytest.raises(RuntimeError, match="does not fit|more denoising slots"):
        step.load_inputs((None, None, _synthetic_targets(30, B, 13, 14)))

    # StepCache: the small batches share one capture (bucket of 2 * num_dn = 40 slots), the dense batch gets its own
    # bucket, same gradient buffer and graph memory pool
    m.load_state_dict(sd)
    cache = dp.StepCache(m, scalar, lambda batch, tgt: (xs, text, tgt), num_dn=20, use_graph=True, warmup=1)
    for b, w in zip(batches, want):
        m.load_state_dict(sd)
        loss = float(cache.run(b))
        assert abs(loss - w[0]) <= 2e-4 * abs(w[0])
    assert list(cache.steps) == [40]
    m.load_state_dict(sd)
    loss = float(cache.run(_synthetic_targets(31, B, 30, 40)))      # up to 2 * 40 = 80 queries -> the 128-slot bucket
    assert list(cache.steps) == [40, 128] and loss == loss
    first, second = cache.steps.values()
    assert first.flat is second.flat and first.graph is not second.graph
    m.load_state_dict(sd)
    assert abs(float(cache.run(batches[1])) - want[1][0]) <= 2e-4 * abs(want[1][0])     # back to the first bucket
