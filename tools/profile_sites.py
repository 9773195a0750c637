"""Where the step's small ATen launches come from: the eager bench step under torch.profiler, device time of the ATen
element-wise / copy / reduce ops grouped (a) by op + input shapes, (b) forward ops by the innermost tamtr_b200 source line,
(c) backward ops by the autograd node that ran them.
    python tools/profile_sites.py [--top 40]"""
import argparse, os, sys, collections, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tamtr_b200 import dp
from tamtr_b200.head import ManbaWorldDecoder
ap = argparse.ArgumentParser(); ap.add_argument("--top", type=int, default=40)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = ManbaWorldDecoder(bench.NC, list(bench.CH), bench.HD, bench.NQ, bench.NDP, bench.NH, bench.NDL, vss=False).to(dev).train()
xs, text = bench.synthetic_inputs(1234, bench.BATCH_PER_GPU, torch.bfloat16)
plan = model.plan_cdn(bench.synthetic_targets(1234, bench.BATCH_PER_GPU))
step = dp.HeadTrainStep(model, bench.surrogate_loss_fn, (xs, text, plan), autocast=torch.bfloat16, use_graph=False)
for _ in range(3): step.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    step.run()
    torch.cuda.synchronize()
evs = list(prof.events())
cpu = [e for e in evs if e.device_type == torch.autograd.DeviceType.CPU]
LIB = ("tamtr::", "nvjet", "cutlass", "cublas", "cudnn")


def kernels_of(e):
    return [k for k in e.kernels] if hasattr(e, "kernels") else []


def is_aten_small(kname):
    return kname.startswith("void at::native") or kname.startswith("at::native") or "Memset" in kname or "Memcpy" in kname or "at_cuda_detail" in kname


# (a) leaf ops (those that directly own kernels) by name + shapes
by_shape = collections.defaultdict(lambda: [0, 0.0])
by_site = collections.defaultdict(lambda: [0, 0.0])
by_node = collections.defaultdict(lambda: [0, 0.0])
nodes = [e for e in cpu if e.name.startswith("autograd::engine::evaluate_function")]
nodes.sort(key=lambda e: e.time_range.start)


def node_of(e):
    for n in nodes:
        if n.thread == e.thread and n.time_range.start <= e.time_range.start and e.time_range.end <= n.time_range.end:
            return n.name.replace("autograd::engine::evaluate_function: ", "")
    return None


for e in cpu:
    ks = [k for k in kernels_of(e) if is_aten_small(k.name)]
    if not ks or any(c for c in e.cpu_children if kernels_of(c)):
        continue
    t = sum(k.duration for k in ks)
    by_shape[(e.name, str(e.input_shapes)[:80])][0] += len(ks); by_shape[(e.name, str(e.input_shapes)[:80])][1] += t
    nd = node_of(e)
    if nd is not None:
        by_node[(nd, e.name)][0] += len(ks); by_node[(nd, e.name)][1] += t
    else:
        site = "?"
        for fr in (e.stack or []):
            if "tamtr_b200/" in fr or "bench.py" in fr:
                site = re.sub(r".*/(tamtr_b200/|bench)", r"\1", fr)[:70]
                break
        by_site[(site, e.name)][0] += len(ks); by_site[(site, e.name)][1] += t

for title, d in (("(a) op + shapes", by_shape), ("(b) forward, by source line", by_site), ("(c) backward, by autograd node", by_node)):
    tot = sum(v[1] for v in d.values()); cnt = sum(v[0] for v in d.values())
    print(f"---- {title}: {cnt} launches, {tot:.0f} us")
    for k, (c, t) in sorted(d.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print(f"  x{c:4d} {t:8.1f} us  {k[0]}  |  {k[1]}")
