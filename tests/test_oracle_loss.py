"""CPU: the loss / matching oracle (oracle/loss_ref.py) against the outputs of the UNMODIFIED reference
(tests/golden/loss.pt, made by oracle/make_goldens_loss.py from ultralytics/models/utils/loss.py + ops.py)."""
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import loss_ref


@pytest.mark.parametrize("name", list(loss_ref.CASES))
def test_oracle_loss_matches_reference(name):
    gold = load_golden("loss")["cases"][name]
    c = loss_ref.make_case(**loss_ref.CASES[name])
    pb = c["pred_bboxes"].clone().requires_grad_()
    ps = c["pred_scores"].clone().requires_grad_()
    kw = {}
    if "dn_meta" in c:
        kw = dict(dn_bboxes=c["dn_bboxes"].clone().requires_grad_(), dn_scores=c["dn_scores"].clone().requires_grad_(),
                  dn_meta=c["dn_meta"])
    loss = loss_ref.rtdetr_detection_loss(pb, ps, c["gt_bboxes"], c["gt_cls"], c["gt_groups"], c["nc"], **kw)
    assert set(loss) == set(gold["loss"])
    for k, v in gold["loss"].items():
        assert abs(float(loss[k]) - v) <= 1e-5 * max(1.0, abs(v)), (k, float(loss[k]), v)
    for l, (img, q, g) in enumerate(gold["matches"]):            # index parity: exact
        mi, mq, mg = loss_ref.hungarian_match(pb[l].detach(), ps[l].detach(), c["gt_bboxes"], c["gt_cls"], c["gt_groups"])
        assert torch.equal(mi, img) and torch.equal(mq, q) and torch.equal(mg, g)
    total = sum(loss.values())
    if total.requires_grad:
        total.backward()
    for t, key in ((pb, "grad_pred_bboxes"), (ps, "grad_pred_scores")):
        if gold[key] is not None:
            assert rel_l2(t.grad, gold[key]) < 1e-5
