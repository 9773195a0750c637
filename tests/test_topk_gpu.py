"""GPU parity of the query-selection top-k kernel (csrc/topk.cu) with `torch.topk(scores, nq, dim=1).indices`
(ultralytics/nn/modules/head.py:1240, :437).  Index work: bit-exact.  The oracle (oracle/topk.py) is a stable descending sort on the CPU --
(score descending, index ascending), the order the kernel documents -- which coincides with torch.topk wherever the scores
are distinct; torch.topk's own values are compared as well."""
import pytest
import torch

from oracle import topk as oracle_topk

pytestmark = pytest.mark.gpu


def _oracle(scores, k):
    return oracle_topk.topk_indices(scores, k)


def _kernel(scores, k):
    """tamtr_topk_rows through the C ABI (ops.topk_rows hands rows that do not fit in shared memory to the library)."""
    from tamtr_b200 import _lib
    scores = scores.cuda().contiguous()
    out = torch.empty(scores.shape[0], k, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().tamtr_topk_rows(scores.data_ptr(), out.data_ptr(), None, scores.shape[0], scores.shape[1], k,
                                          _lib.stream_ptr(scores.device)), "topk_rows")
    return out


@pytest.mark.parametrize("B,n,k", [(16, 33600, 300),      # TAMTR.yaml pyramid at 640^2, S-yaml queries (keys in shared memory)
                                   (2, 8400, 300),        # RTDETRDecoder config 1
                                   (3, 134400, 900),      # 1280^2 inference (row re-read from global memory)
                                   (4, 100, 100),         # k == n
                                   (5, 37, 5), (1, 5000, 1), (2, 1, 1), (2, 4097, 4096)])
def test_topk_rows_matches_library_on_distinct_scores(cuda_lib, B, n, k):
    from tamtr_b200 import ops
    g = torch.Generator().manual_seed(n * 31 + k)
    scores = (torch.randn(B, n, generator=g) * 3.0 - 1.0)
    idx = _kernel(scores, k)
    ref = _oracle(scores, k)
    assert torch.equal(idx.cpu(), ref)
    via_ops = ops.topk_rows(scores.cuda(), k)
    assert via_ops.dtype == torch.int64 and via_ops.shape == (B, k) and torch.equal(via_ops.cpu(), ref)
    lib_vals = torch.topk(scores.cuda(), k, dim=1).values
    assert torch.equal(torch.gather(scores.cuda(), 1, idx), lib_vals)


@pytest.mark.parametrize("n,k", [(33600, 300), (134400, 900), (513, 64)])
def test_topk_rows_ties_and_special_values(cuda_lib, n, k):
    """Heavily tied rows (quantised scores; a constant row; masked tokens sharing one score as in head.py:1229 `valid * feats`),
    infinities, denormals: winners among equal scores are the lowest indices, in index order."""
    from tamtr_b200 import ops
    g = torch.Generator().manual_seed(n + k)
    rows = [torch.round(torch.randn(n, generator=g) * 4) / 4,                  # ~60 distinct values
            torch.full((n,), 0.125),                                           # constant row
            torch.where(torch.rand(n, generator=g) < 0.7, torch.tensor(-2.5), torch.randn(n, generator=g)),
            torch.randn(n, generator=g) * 1e-42,                               # denormals (both signs)
            torch.randn(n, generator=g)]
    rows[4][torch.randperm(n, generator=g)[:7]] = float("inf")
    rows[4][torch.randperm(n, generator=g)[:n // 2]] = float("-inf")
    scores = torch.stack(rows)                 # row 0 holds +0.0 and -0.0: equal scores, as for torch
    idx = _kernel(scores, k)
    assert torch.equal(idx.cpu(), _oracle(scores, k))
    keep = [0, 1, 2, 4]                        # (whether the library's float compares flush denormals is its own business)
    assert torch.equal(torch.gather(scores.cuda(), 1, idx)[keep], torch.topk(scores.cuda()[keep], k, dim=1).values)


def test_topk_rows_values_output_and_argument_checks(cuda_lib):
    from tamtr_b200 import _lib
    L = _lib.lib()
    scores = torch.randn(3, 777, generator=torch.Generator().manual_seed(5)).cuda()
    idx = torch.empty(3, 20, dtype=torch.int64, device="cuda")
    val = torch.empty(3, 20, dtype=torch.float32, device="cuda")
    st = _lib.stream_ptr(scores.device)
    assert L.tamtr_topk_rows(scores.data_ptr(), idx.data_ptr(), val.data_ptr(), 3, 777, 20, st) == 0
    ref = torch.topk(scores, 20, dim=1)
    assert torch.equal(val, ref.values) and torch.equal(idx, ref.indices)
    assert L.tamtr_topk_rows_supported(777, 778) == 0 and L.tamtr_topk_rows_supported(10000, 4097) == 0
    assert L.tamtr_topk_rows_supported(33600, 300) == 2 and L.tamtr_topk_rows_supported(134400, 900) == 1
    assert L.tamtr_topk_rows(scores.data_ptr(), idx.data_ptr(), None, 3, 10, 20, st) != 0          # k > n
    assert L.tamtr_topk_rows(None, idx.data_ptr(), None, 3, 777, 20, st) != 0
    # non-fp32 / CPU scores take the library call
    from tamtr_b200 import ops
    half = scores.bfloat16()
    assert torch.equal(ops.topk_rows(half, 9), torch.topk(half, 9, dim=1).indices)
    assert torch.equal(ops.topk_rows(scores.cpu(), 9), torch.topk(scores.cpu(), 9, dim=1).indices)


def test_topk_rows_in_a_cuda_graph(cuda_lib):
    from tamtr_b200 import ops
    scores = torch.randn(4, 33600, generator=torch.Generator().manual_seed(9)).cuda()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ops.topk_rows(scores, 300)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = ops.topk_rows(scores, 300)
    scores.copy_(torch.randn(4, 33600, generator=torch.Generator().manual_seed(10)))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), _oracle(scores, 300))
