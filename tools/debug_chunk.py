import sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamtr_b200
from tamtr_b200 import vss
from tamtr_b200.vss import VSSBlock
blk = VSSBlock(hidden_dim=128, drop_path=0.0).cuda().eval()
x = torch.randn(1, 320, 320, 128, device="cuda")
orig = vss._SelectiveScanFn.forward
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    b = tamtr_b200.launch_count(); blk(x); print("launches", tamtr_b200.launch_count() - b)
def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    c.record(); torch.cuda.synchronize()
    return a.elapsed_time(c) / n
def run():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        return blk(x)
print("chunked", timed(run))
vss.CHUNKED_INFERENCE = False
print("plain", timed(run))
from torch.profiler import profile, ProfilerActivity
vss.CHUNKED_INFERENCE = True
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=60))
