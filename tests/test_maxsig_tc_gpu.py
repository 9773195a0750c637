"""GPU parity of the tcgen05/TMEM/TMA max-sigmoid gate against the CPU oracle arithmetic (fp32 einsum on the same
bf16-rounded operands) and against the exact CUDA-core kernel."""
import pytest
import torch

from helpers import rel_l2
from oracle import seeding

pytestmark = pytest.mark.gpu

CASES = [  # B, nh, H, W, N
    (2, 8, 40, 40, 10),      # TAMTR.yaml layer 16/40: C=256, 12.5 pixel tiles -> TMA zero fill on the last one
    (1, 2, 160, 160, 10),    # TAMTR.yaml layer 32: C=64
    (2, 4, 80, 80, 10),      # C=128
    (2, 8, 80, 80, 80),      # config-3 sweep corner: 80 text tokens
    (1, 8, 20, 20, 17),      # N not a multiple of 16 -> padded columns must never win the max
    (1, 1, 8, 8, 1),
]


def _reference(xb, g, bias, nh):
    B, C, H, W = xb.shape
    hc = C // nh
    e = xb.float().view(B, nh, hc, H, W)
    gb = g.bfloat16().float()                      # the tensor-core path feeds the guide as bf16
    logits = torch.einsum("bmchw,bnmc->bmhwn", e, gb)
    best, arg = logits.max(-1)
    return (best / hc ** 0.5 + bias[None, :, None, None]).sigmoid(), arg, logits


@pytest.mark.parametrize("B,nh,H,W,N", CASES)
def test_tensor_core_gate_matches_oracle(cuda_lib, B, nh, H, W, N):
    from tamtr_b200 import ops
    hc = 32
    x = seeding.seeded_tensor(N + H, "x", (B, nh * hc, H, W))
    g = seeding.seeded_tensor(N + H, "g", (B, N, nh, hc)) * 0.4 - 0.3      # mostly negative logits: padding would win
    bias = seeding.seeded_tensor(N + H, "b", (nh,))
    xb = x.bfloat16()
    ref, arg_ref, logits = _reference(xb, g, bias, nh)
    aw = ops.max_sigmoid_gate(xb.cuda(), g.cuda(), bias.cuda(), nh, use_tensor_cores=True)
    torch.cuda.synchronize()
    assert aw.shape == (B, nh, H, W) and aw.dtype == torch.float32
    assert rel_l2(aw, ref) < 1e-5, rel_l2(aw, ref)           # same operands, fp32 accumulate: only summation order
    exact = ops.max_sigmoid_gate(xb.cuda(), g.cuda(), bias.cuda(), nh, use_tensor_cores=False)
    assert rel_l2(aw, exact) < 2e-2                           # bf16 guide vs fp32 guide: the bf16 budget


def test_tensor_core_gate_backward(cuda_lib):
    """The backward (shared CUDA-core kernel) consumes the arg-max written by the tensor-core forward.  Reference:
    autograd of the fp32 einsum on the same bf16-rounded operands (so that the arg-max -- which routes the whole
    gradient -- is decided on the same numbers)."""
    from tamtr_b200 import ops
    B, nh, hc, H, W, N = 2, 8, 32, 40, 40, 10
    x = seeding.seeded_tensor(3, "x", (B, nh * hc, H, W)).bfloat16()
    g = (seeding.seeded_tensor(3, "g", (B, N, nh, hc)) * 0.3).bfloat16().float()
    bias = seeding.seeded_tensor(3, "b", (nh,))
    go = seeding.seeded_tensor(3, "go", (B, nh, H, W))
    xr, gr, br = x.float().requires_grad_(), g.clone().requires_grad_(), bias.clone().requires_grad_()
    e = xr.view(B, nh, hc, H, W)
    ref = (torch.einsum("bmchw,bnmc->bmhwn", e, gr).max(-1)[0] / hc ** 0.5 + br[None, :, None, None]).sigmoid()
    ref.backward(go)
    xc, gc, bc = x.cuda().requires_grad_(), g.cuda().requires_grad_(), bias.cuda().requires_grad_()
    aw = ops.max_sigmoid_gate(xc, gc, bc, nh, use_tensor_cores=True)
    aw.backward(go.cuda())
    assert rel_l2(aw, ref) < 1e-5
    assert rel_l2(xc.grad, xr.grad) < 2e-2          # grad_x is stored in bf16
    assert rel_l2(gc.grad, gr.grad) < 1e-4 and rel_l2(bc.grad, br.grad) < 1e-4
