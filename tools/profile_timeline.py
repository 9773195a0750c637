"""Timeline of one graph-replayed bench step (CUPTI through torch.profiler): span, kernel time on the main chain, idle gaps,
where the forked zero fill sits, the largest gaps and what surrounds them.
    python tools/profile_timeline.py [--steps 4]"""
import argparse, os, sys, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from tamtr_b200 import dp
from tamtr_b200.head import ManbaWorldDecoder
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=4); ap.add_argument("--gaps", type=int, default=12)
ap.add_argument("--device-targets", action="store_true", help="ground truth as loss.DeviceTargets (bench.py's step) instead of a host plan")
ap.add_argument("--no-profile", action="store_true")
ap.add_argument("--loops", type=int, default=4, help="timed loops of 20 replays (many: does the mode change while the GPU stays busy?)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
model = ManbaWorldDecoder(bench.NC, list(bench.CH), bench.HD, bench.NQ, bench.NDP, bench.NH, bench.NDL, vss=False).to(dev).train()
xs, text = bench.synthetic_inputs(1234, bench.BATCH_PER_GPU, torch.bfloat16)
plan = bench.device_targets(bench.synthetic_targets(1234, bench.BATCH_PER_GPU), dev) if args.device_targets \
    else model.plan_cdn(bench.synthetic_targets(1234, bench.BATCH_PER_GPU))
step = dp.HeadTrainStep(model, bench.surrogate_loss_fn, (xs, text, plan), autocast=torch.bfloat16, use_graph=True)
for _ in range(5): step.run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(args.loops):
    a.record()
    for _ in range(20): step.run()
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 20)
    if args.loops > 8:
        import datetime
        print(datetime.datetime.now().strftime('%H:%M:%S.%f')[:-3], f'{ts[-1]:.3f}', flush=True)
print("events: " + " ".join(f"{t:.4f}" if args.loops <= 8 else f"{t:.2f}" for t in ts) + f" ms/step (20 replays back to back, {args.loops} times)")
if args.no_profile:
    sys.exit(0)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.steps):
        step.run()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
n = len(evs) // args.steps
for s in range(1, args.steps):            # skip the first replay under the profiler
    one = evs[s * n:(s + 1) * n]
    t0 = one[0].time_range.start
    fill = [e for e in one if "zero_fill_bg" in e.name]
    main = [e for e in one if "zero_fill_bg" not in e.name]
    span = max(e.time_range.end for e in one) - t0
    busy = 0.0; cur_end = t0; gaps = []
    for i, e in enumerate(main):
        st, en = e.time_range.start, e.time_range.end
        if st > cur_end:
            gaps.append((st - cur_end, i))
        busy += max(0.0, en - max(st, cur_end)); cur_end = max(cur_end, en)
    nxt = evs[(s + 1) * n].time_range.start - t0 if (s + 1) * n < len(evs) else float("nan")
    print(f"step {s}: span {span:.1f} us, start-to-next-start {nxt:.1f} us, main-chain busy {busy:.1f} us, idle {span - busy:.1f} us in {len(gaps)} gaps, kernels {len(one)}")
    for f in fill:
        fs, fe = f.time_range.start - t0, f.time_range.end - t0
        during = [e for e in main if e.time_range.end > f.time_range.start and e.time_range.start < f.time_range.end]
        print(f"   zero fill: {fs:.1f} .. {fe:.1f} us ({fe - fs:.1f} us), beside {len(during)} main-chain kernels, "
              f"first: {re.sub(r'<.*', '', during[0].name)[:50] if during else '-'} last: {re.sub(r'<.*', '', during[-1].name)[:50] if during else '-'}")
        nb = [e for e in main if 'msda_bwd' in e.name]
        if nb:
            print(f"   first sampler backward starts at {nb[0].time_range.start - t0:.1f} us; last tok_project ends at "
                  f"{max(e.time_range.end for e in main if 'tok_project' in e.name) - t0:.1f} us")
    if s == args.steps - 1:
        print(f"   largest gaps:")
        for g, i in sorted(gaps, reverse=True)[:args.gaps]:
            print(f"     {g:7.1f} us at {main[i].time_range.start - t0:8.1f} us  after {re.sub(r'<.*', '', main[i-1].name)[:45]:45s} before {re.sub(r'<.*', '', main[i].name)[:45]}")
        hist = {}
        for g, _ in gaps:
            k = "<1" if g < 1 else "<2" if g < 2 else "<4" if g < 4 else "<8" if g < 8 else ">=8"
            hist.setdefault(k, [0, 0.0]); hist[k][0] += 1; hist[k][1] += g
        print("   gap histogram (us): " + ", ".join(f"{k}: {c} gaps {t:.0f} us" for k, (c, t) in sorted(hist.items())))
