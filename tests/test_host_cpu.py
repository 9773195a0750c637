"""Host-side logic that needs no GPU: routing of ops.topk_rows, the deferred fork of the gradient arena's zero fill, and
the control logic of bench.py's steady-state probe (fed with loop times recorded on a B200: profiles/step_mode_r2.txt)."""
import importlib.util
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_topk_rows_on_cpu_is_the_library_call():
    from tamtr_b200 import ops
    scores = torch.randn(3, 50, generator=torch.Generator().manual_seed(0))
    assert torch.equal(ops.topk_rows(scores, 7), torch.topk(scores, 7, dim=1).indices)
    # query selection of the reference (head.py:1240): indices into dim 1, best first
    assert torch.equal(torch.gather(scores, 1, ops.topk_rows(scores, 7)), torch.sort(scores, 1, descending=True).values[:, :7])


def test_arena_prefill_request_is_inert_without_cuda():
    from tamtr_b200 import ops
    arena = ops.ValueArena()
    arena.defer = True
    arena.prefill((2, 10, 8), torch.bfloat16, torch.device("cpu"))        # CPU: neither forks nor records a request
    assert arena.buf is None and arena._pending is None
    arena.start_prefill()                                                 # nothing pending: no-op, and the flag is spent
    assert arena.buf is None and arena.defer is False
    arena.join()
    buf, written = arena.take()
    assert buf is None and written == set()


def _bench():
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_steady_state_probe_follows_the_recorded_change():
    probe = _bench().steady_state_probe
    steps = 20
    # call M2, process 2: 46 loops at ~4.20 ms, one at 4.28 ms, then 4.02 ms
    recorded = [4.21] * 4 + [4.20] * 42 + [4.28] + [4.02] * 16 + [4.11] + [4.02] * 7
    it = iter(t * 1e-3 * steps for t in recorded)
    r = probe(lambda: next(it), steps, 16)
    assert abs(r["ms_per_step_first"] - 4.21) < 1e-9 and abs(r["ms_per_step_last"] - 4.02) < 1e-9
    assert r["loops"] == 50 and 3.9 < r["changed_after_s"] < 4.1          # three settled loops after the change, then stop
    assert abs(r["value_last"] - 16 / 4.02e-3) < 1e-6
    # a process that never changes state: stops at the time budget and reports the same state
    it = iter([4.21e-3 * steps] * 10000)
    r = probe(lambda: next(it), steps, 16, budget_s=2.0)
    assert r["changed_after_s"] is None and abs(r["ms_per_step_last"] - 4.21) < 1e-9 and 2.0 <= r["gpu_seconds"] < 2.1
    # one noisy fast loop does not count as the change having settled
    recorded = [4.21] * 3 + [4.00] + [4.21] * 30
    it = iter(t * 1e-3 * steps for t in recorded + [4.21] * 10000)
    r = probe(lambda: next(it), steps, 16, budget_s=1.0)
    assert abs(r["ms_per_step_last"] - 4.21) < 1e-9
