"""Kernel-time breakdown of the bench step with torch.profiler (CUPTI), on the CUDA-graph replay.
    python tools/profile_step.py [--eager] [--top 40]"""
import argparse, os, sys, collections, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tamtr_b200 import dp
from tamtr_b200.head import ManbaWorldDecoder
ap = argparse.ArgumentParser(); ap.add_argument("--eager", action="store_true"); ap.add_argument("--top", type=int, default=45)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = ManbaWorldDecoder(bench.NC, list(bench.CH), bench.HD, bench.NQ, bench.NDP, bench.NH, bench.NDL, vss=False).to(dev).train()
xs, text = bench.synthetic_inputs(1234, bench.BATCH_PER_GPU, torch.bfloat16)
plan = model.plan_cdn(bench.synthetic_targets(1234, bench.BATCH_PER_GPU))
step = dp.HeadTrainStep(model, bench.surrogate_loss_fn, (xs, text, plan), autocast=torch.bfloat16, use_graph=not args.eager)
for _ in range(3): step.run()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.steps): step.run()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"<.*", "", ev.name)[:90]
        agg[name][0] += 1; agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"steps={args.steps} total kernel time/step = {tot/args.steps/1e3:.3f} ms, kernels/step = {sum(v[0] for v in agg.values())/args.steps:.0f}")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:args.top]:
    print(f"{t/args.steps:10.1f} us {100*t/tot:5.1f}%  x{c/args.steps:6.1f}  {name}")

print("---- individual launches > 40 us (one step)")
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // args.steps
for ev in evs[:n]:
    if ev.device_time > 40:
        print(f"{ev.device_time:9.1f} us  {ev.name[:150]}")
print("---- small kernels (< 12 us) by name, one step")
small = collections.defaultdict(lambda: [0, 0.0])
for ev in evs[:n]:
    if ev.device_time <= 12:
        nm = re.sub(r"\(.*", "", ev.name)[:110]
        small[nm][0] += 1; small[nm][1] += ev.device_time
tot_s = sum(v[1] for v in small.values()); cnt_s = sum(v[0] for v in small.values())
print(f"small kernels: {cnt_s} launches, {tot_s/1e3:.3f} ms")
for nm, (c, t) in sorted(small.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  x{c:4d} {t:8.1f} us  {nm}")
