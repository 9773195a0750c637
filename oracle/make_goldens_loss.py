"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/loss.pt by running the UNMODIFIED reference's
RTDETRDetectionLoss / HungarianMatcher (ultralytics/models/utils/loss.py, ops.py; configured as nn/tasks.py:578)
on the seeded cases of oracle/loss_ref.py:CASES, in this container (CPU).

    python -m oracle.make_goldens_loss
"""
import importlib
import os

import torch

from oracle import loss_ref, reference_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "loss.pt")


def main():
    reference_loader.load()
    L = importlib.import_module("ultralytics.models.utils.loss")
    out = {"source": "reference RTDETRDetectionLoss(nc, use_vfl=True) + HungarianMatcher (scipy) on CPU, torch "
                     + torch.__version__, "cases": {}}
    for name, spec in loss_ref.CASES.items():
        c = loss_ref.make_case(**spec)
        crit = L.RTDETRDetectionLoss(nc=c["nc"], use_vfl=True, use_sl=False, use_emasl=False, use_svfl=False,
                                     use_emasvfl=False)
        pb = c["pred_bboxes"].clone().requires_grad_()
        ps = c["pred_scores"].clone().requires_grad_()
        batch = {"cls": c["gt_cls"], "bboxes": c["gt_bboxes"], "gt_groups": c["gt_groups"]}
        kw = {}
        if "dn_meta" in c:
            db = c["dn_bboxes"].clone().requires_grad_()
            ds = c["dn_scores"].clone().requires_grad_()
            kw = dict(dn_bboxes=db, dn_scores=ds, dn_meta=c["dn_meta"])
        loss = crit((pb, ps), batch, **kw)
        total = sum(loss.values())
        total.backward()
        matches = []
        for l in range(pb.shape[0]):
            m = crit.matcher(pb[l].detach(), ps[l].detach(), c["gt_bboxes"], c["gt_cls"], c["gt_groups"])
            matches.append((torch.cat([torch.full_like(i, b) for b, (i, _) in enumerate(m)]),
                            torch.cat([i for i, _ in m]), torch.cat([j for _, j in m])))
        out["cases"][name] = {
            "loss": {k: float(v.detach()) for k, v in loss.items()},
            "total": float(total.detach()),
            "matches": matches,
            "grad_pred_bboxes": None if pb.grad is None else pb.grad.clone(),
            "grad_pred_scores": None if ps.grad is None else ps.grad.clone(),
            "grad_dn_bboxes": kw["dn_bboxes"].grad.clone() if kw else None,
            "grad_dn_scores": kw["dn_scores"].grad.clone() if kw else None,
        }
        print(name, {k: round(v, 5) for k, v in out["cases"][name]["loss"].items()})
    torch.save(out, OUT)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
