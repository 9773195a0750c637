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


def test_gradient_buckets_partition_the_flat_buffer():
    """FlatGrads.set_buckets / bucket_slice / pack_bucket: the buckets tile the padded flat buffer without gaps and packing
    them one by one equals packing everything at once (what the overlapped all-reduce relies on)."""
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(7, 13), torch.nn.Linear(13, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    net(torch.randn(4, 7)).sum().backward()
    net[2].bias.grad = None                                      # a parameter without a gradient is packed as zeros
    flat = dp.FlatGrads(list(net.parameters()))
    assert all(o % 4 == 0 for o in flat.offsets) and flat.flat.numel() % 4 == 0
    whole = flat.gather().clone()
    for n in (1, 2, 3, 8):
        flat.set_buckets(n)
        assert flat.buckets[0][0] == 0 and flat.buckets[-1][1] == len(flat.params)
        assert all(a[1] == b[0] for a, b in zip(flat.buckets, flat.buckets[1:]))
        assert sum(flat.bucket_slice(b).numel() for b in range(len(flat.buckets))) == flat.flat.numel()
        flat.flat.fill_(7.0)
        for b in range(len(flat.buckets)):
            flat.pack_bucket(b)
        pad = torch.ones_like(whole, dtype=torch.bool)
        for o, p in zip(flat.offsets, flat.params):
            pad[o:o + p.numel()] = False
        assert torch.equal(flat.flat[~pad], whole[~pad])


def test_load_inputs_refuses_what_it_cannot_update():
    """A captured step bakes non-tensor arguments in: replacing one must fail loudly (a new CdnPlan used to be
    ignored silently)."""
    net = torch.nn.Linear(4, 2)

    class Plan:
        def to(self, device):
            return self

        def materialize(self, w):
            return None

    step = dp.HeadTrainStep(torch.nn.Sequential(net), lambda o: o.sum(), (torch.zeros(3, 4),), use_graph=False)
    step.static.append(Plan())
    step.load_inputs((torch.ones(3, 4), None))                   # None keeps the captured object
    assert float(step.static[0].sum()) == 12.0
    with pytest.raises(RuntimeError, match="cannot be updated"):
        step.load_inputs((torch.ones(3, 4), Plan()))


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
