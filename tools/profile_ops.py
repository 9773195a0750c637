"""Which ATen ops (and input shapes) the eager head training step spends device time in: torch.profiler, self CUDA time
grouped by op and shape.  `python tools/profile_ops.py [--vss] [--top 60]`"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from tamtr_b200 import dp  # noqa: E402
from tamtr_b200.head import ManbaWorldDecoder  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vss", action="store_true")
ap.add_argument("--top", type=int, default=60)
ap.add_argument("--loss", default="surrogate")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.manual_seed(1234)
model = ManbaWorldDecoder(bench.NC, list(bench.CH), bench.HD, bench.NQ, bench.NDP, bench.NH, bench.NDL, vss=args.vss).to(dev).train()
if args.vss:
    for blk in model.VSSBlocks:
        blk.drop_path.drop_prob = 0.0
xs, text = bench.synthetic_inputs(1234, bench.BATCH_PER_GPU, torch.bfloat16)
plan = bench.device_targets(bench.synthetic_targets(1234, bench.BATCH_PER_GPU), dev)
loss_fn = bench.make_detection_loss() if args.loss == "detection" else bench.surrogate_loss_fn
step = dp.HeadTrainStep(model, loss_fn, ([x.to(dev) for x in xs], text.to(dev), plan), autocast=torch.bfloat16, use_graph=False)
for _ in range(3):
    step.run()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True,
             experimental_config=torch._C._profiler._ExperimentalConfig(verbose=True)) as prof:
    step.run()
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.self_device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"self device time of one eager step: {tot / 1e3:.3f} ms over {sum(e.count for e in rows)} op calls")
for e in rows[:args.top]:
    shapes = str(e.input_shapes)[:110]
    print(f"{e.self_device_time_total:9.1f} us  x{e.count:4d}  {e.key[:44]:44s} {shapes}")

# ---- small elementwise / copy ops by the product-code line that issued them
import collections  # noqa: E402
by_site = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.self_device_time_total <= 0 or ev.device_type != torch.autograd.DeviceType.CPU:
        continue
    if not ev.name.startswith("aten::") or ev.name in ("aten::mm", "aten::addmm", "aten::bmm"):
        continue
    frames = [f for f in (ev.stack or []) if "tamtr_b200/" in f or "bench.py" in f]
    site = " < ".join(f.split("tamtr_b200/")[-1].split("/root/repo/")[-1][:48] for f in frames[:2]) or "?"
    shape = str(ev.input_shapes[0])[:22] if ev.input_shapes else ""
    by_site[(ev.name, site, shape)][0] += 1
    by_site[(ev.name, site, shape)][1] += ev.self_device_time_total
print("---- aten elementwise / copy ops by call site")
for (name, site, shape), (c, t) in sorted(by_site.items(), key=lambda kv: -kv[1][1])[:110]:
    print(f"{t:8.1f} us x{c:4d}  {name[6:]:18s} {shape:22s} {site}")
