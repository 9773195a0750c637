"""N>1 path on CPU: world_size-2 gloo processes exercise the sharding, the flat-gradient all-reduce and the
max-over-ranks timing reduction of tamtr_b200/dp.py (the CUDA graph / NCCL parts need GPUs)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tamtr_b200 import dp


def test_shard_indices_partition_the_batch():
    for n in (1, 7, 16, 64, 65):
        for ws in (1, 2, 3, 8):
            parts = [list(dp.shard_indices(n, r, ws)) for r in range(ws)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        torch.manual_seed(0)                                   # same weights on every rank (DDP's initial broadcast)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        data = torch.arange(12 * 8, dtype=torch.float32).view(12, 8) / 50.0
        mine = data[list(dp.shard_indices(12, rank, ws))]      # images are the independent units: no collective
        step = dp.HeadTrainStep(model, lambda o: o.square().mean(), (mine,), use_graph=False)
        loss = step.run()
        flat = step.flat.flat.clone()
        # reference: average of the per-rank gradients computed independently
        grads = []
        for r in range(ws):
            m2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
            m2.load_state_dict(model.state_dict())
            m2(data[list(dp.shard_indices(12, r, ws))]).square().mean().backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in m2.parameters()]))
        expect = torch.stack(grads).mean(0)
        ok_grad = torch.allclose(flat, expect, atol=1e-6)
        step.flat.scatter()
        views_ok = all(p.grad.data_ptr() >= step.flat.flat.data_ptr() for p in model.parameters())
        t = dp.max_over_ranks(1.0 + rank, torch.device("cpu"))
        second = step.run()                                    # grads are re-zeroed each step, not accumulated
        ok_second = torch.allclose(step.flat.flat, expect, atol=1e-6)
        out.put((rank, ok_grad, views_ok, t, ok_second, float(loss)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_all_reduce_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=100) for _ in procs)
    for p in procs:
        p.join(30)
    for rank, ok_grad, views_ok, t, ok_second, loss in res:
        assert ok_grad and views_ok and ok_second
        assert t == 2.0                                         # max over ranks, identical on every rank
