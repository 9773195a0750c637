"""Device time of the detection loss at the bench shapes: cost matrices, the assignment kernel, loss forward+backward."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loss_ref  # noqa: E402
from tamtr_b200.loss import RTDETRDetectionLoss  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    g = torch.Generator().manual_seed(1234)
    groups = [100] + [int(torch.randint(20, 101, (1,), generator=g)) for _ in range(15)]
    c = loss_ref.make_case(seed=7, n_layers=4, bs=16, nq=100, nc=10, gt_groups=groups, dn_groups=1)
    crit = RTDETRDetectionLoss(nc=10, use_vfl=True)
    pb = c["pred_bboxes"].cuda().requires_grad_()
    ps = c["pred_scores"].cuda().requires_grad_()
    batch = {"cls": c["gt_cls"].cuda(), "bboxes": c["gt_bboxes"].cuda(), "gt_groups": groups}
    kw = dict(dn_bboxes=c["dn_bboxes"].cuda().requires_grad_(), dn_scores=c["dn_scores"].cuda().requires_grad_(),
              dn_meta=c["dn_meta"])
    from tamtr_b200.loss import DeviceTargets
    tgt = DeviceTargets.from_batch(batch, pb.device)
    m = crit.matcher
    print("cost + assignment %7.1f us  (4 layers x 16 images, %d gts; two launches, no host sync)"
          % (timed(lambda: m.match_padded(pb, ps, tgt)), sum(groups)))

    def full():
        pb.grad = ps.grad = None
        sum(crit((pb, ps), batch, **kw).values()).backward()
    print("loss fwd+bwd     %8.1f us  (eager)" % timed(full))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        full()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        full()
    print("loss fwd+bwd     %8.1f us  (graph replay)" % timed(graph.replay))


if __name__ == "__main__":
    main()
