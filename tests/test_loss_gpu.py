"""GPU: the device-side Hungarian matching (csrc/assign.cu) and the mirrored detection loss (tamtr_b200/loss.py) against
(i) the reference's own outputs (tests/golden/loss.pt), (ii) scipy.optimize.linear_sum_assignment -- the call the
reference makes at ultralytics/models/utils/ops.py:117 -- on random, integer (tie-heavy), constant and ragged cost
matrices.  Index parity is exact."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment as scipy_lsa

from helpers import load_golden, rel_l2
from oracle import loss_ref

pytestmark = pytest.mark.gpu


def _check_against_scipy(C, gt_groups):
    from tamtr_b200.loss import linear_sum_assignment
    img, q, g = linear_sum_assignment(C.cuda(), gt_groups)
    img, q, g = img.cpu(), q.cpu(), g.cpu()
    n_layers, bs, nq, _ = C.shape
    for l in range(n_layers):
        ei, eq, eg, start = [], [], [], 0
        for b, n in enumerate(gt_groups):
            i, j = scipy_lsa(C[l, b, :, start:start + n].numpy())
            ei += [b] * len(i)
            eq += list(i)
            eg += list(j + start)
            start += n
        assert img.tolist() == ei
        assert q[l].tolist() == eq, f"layer {l}: query indices differ"
        assert g[l].tolist() == eg, f"layer {l}: gt indices differ"


@pytest.mark.parametrize("kind", ["random", "integer_ties", "constant", "ragged_big", "one_gt", "wide"])
def test_assignment_equals_scipy(cuda_lib, kind):
    gen = torch.Generator().manual_seed(11)
    if kind == "random":
        groups, nq, L = [23, 0, 57, 100, 1], 100, 3
        C = torch.randn(L, len(groups), nq, sum(groups), generator=gen)
    elif kind == "integer_ties":        # small integer costs: many equal path costs, SciPy's tie rules decide
        groups, nq, L = [17, 40, 64], 64, 2
        C = torch.randint(0, 4, (L, len(groups), nq, sum(groups)), generator=gen).float()
    elif kind == "constant":            # SciPy #11602: a constant matrix must give the identity assignment
        groups, nq, L = [30, 50], 50, 1
        C = torch.full((L, len(groups), nq, sum(groups)), 3.5)
    elif kind == "ragged_big":          # more gts than queries in one image; sub-matrix too big for shared memory
        groups, nq, L = [350, 12], 300, 1
        C = torch.randn(L, len(groups), nq, sum(groups), generator=gen)
    elif kind == "one_gt":
        groups, nq, L = [1, 1], 7, 1
        C = torch.randn(L, len(groups), nq, sum(groups), generator=gen)
    else:                               # 900 queries (BASELINE.json config 5)
        groups, nq, L = [120, 64], 900, 1
        C = torch.randn(L, len(groups), nq, sum(groups), generator=gen)
    _check_against_scipy(C, groups)


@pytest.mark.parametrize("name", list(loss_ref.CASES))
def test_loss_matches_reference(cuda_lib, name):
    from tamtr_b200.loss import RTDETRDetectionLoss
    gold = load_golden("loss")["cases"][name]
    c = loss_ref.make_case(**loss_ref.CASES[name])
    crit = RTDETRDetectionLoss(nc=c["nc"], use_vfl=True)
    pb = c["pred_bboxes"].cuda().requires_grad_()
    ps = c["pred_scores"].cuda().requires_grad_()
    batch = {"cls": c["gt_cls"].cuda(), "bboxes": c["gt_bboxes"].cuda(), "gt_groups": c["gt_groups"]}
    kw = {}
    if "dn_meta" in c:
        kw = dict(dn_bboxes=c["dn_bboxes"].cuda().requires_grad_(), dn_scores=c["dn_scores"].cuda().requires_grad_(),
                  dn_meta=c["dn_meta"])
    loss = crit((pb, ps), batch, **kw)
    assert set(loss) == set(gold["loss"])
    for k, v in gold["loss"].items():
        assert abs(float(loss[k]) - v) <= 2e-5 * max(1.0, abs(v)), (k, float(loss[k]), v)
    if sum(c["gt_groups"]):
        img, q, g = crit.matcher.match_layers(pb.detach(), ps.detach(), batch["bboxes"], batch["cls"], c["gt_groups"])
        for l, (gi, gq, gg) in enumerate(gold["matches"]):        # the reference's matches, exactly
            assert torch.equal(img.cpu(), gi) and torch.equal(q[l].cpu(), gq) and torch.equal(g[l].cpu(), gg)
        per_image = crit.matcher(pb[0].detach(), ps[0].detach(), batch["bboxes"], batch["cls"], c["gt_groups"])
        assert len(per_image) == len(c["gt_groups"])               # the reference's return convention
        assert torch.equal(torch.cat([i for i, _ in per_image]).cpu(), gold["matches"][0][1])
    total = sum(loss.values())
    total.backward()
    for t, key in ((pb, "grad_pred_bboxes"), (ps, "grad_pred_scores")):
        if gold[key] is not None:
            assert rel_l2(t.grad, gold[key]) < 2e-5
    if kw:
        assert rel_l2(kw["dn_bboxes"].grad, gold["grad_dn_bboxes"]) < 2e-5
        assert rel_l2(kw["dn_scores"].grad, gold["grad_dn_scores"]) < 2e-5


def test_loss_step_is_graph_capturable(cuda_lib):
    """No host synchronisation inside the loss: forward + backward captured into a CUDA graph and replayed on new
    predictions gives the same numbers as eager."""
    from tamtr_b200.loss import RTDETRDetectionLoss
    c = loss_ref.make_case(**loss_ref.CASES["visdrone_like"])
    crit = RTDETRDetectionLoss(nc=c["nc"], use_vfl=True)
    batch = {"cls": c["gt_cls"].cuda(), "bboxes": c["gt_bboxes"].cuda(), "gt_groups": c["gt_groups"]}
    pb = c["pred_bboxes"].cuda().requires_grad_()
    ps = c["pred_scores"].cuda().requires_grad_()
    out = torch.zeros((), device="cuda")

    def step():
        pb.grad = ps.grad = None
        total = sum(crit((pb, ps), batch).values())
        total.backward()
        out.copy_(total.detach())

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(side)
    eager, eager_grad = float(out), pb.grad.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    with torch.no_grad():
        pb.mul_(0.97)
    graph.replay()
    replay1 = float(out)
    with torch.no_grad():
        pb.div_(0.97)
    graph.replay()
    assert abs(float(out) - eager) < 1e-4 * abs(eager) and replay1 != float(out)
    assert rel_l2(pb.grad, eager_grad) < 1e-5
