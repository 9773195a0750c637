import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_step_gpu import _setup, _loss, _grads
from tamtr_b200 import dp, ops
mode = sys.argv[1]
m, xs, text, plan = _setup()
if mode == "eager_first":
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = m([x.cuda() for x in xs], text.cuda(), plan.to("cuda"))
    _loss(out).backward()
    print("eager done")
try:
    step = dp.HeadTrainStep(m, _loss, (xs, text, plan), autocast=torch.bfloat16, use_graph=True, fused_param_cast=(mode != "nofuse"), warmup=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    print(mode, "capture ok", step.run().item())
except Exception as e:
    print(mode, "capture FAILED", str(e)[:200])
