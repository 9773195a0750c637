"""CPU: the VSSBlock / SS2D oracle (oracle/vss_ref.py) against the UNMODIFIED reference VSSBlock run with the scan
plugged in (tests/golden/vss.pt, oracle/make_goldens_vss.py), the scan restatement's own gradients (fp64 gradcheck), and
the product module's state_dict layout against the reference's."""
import pytest
import torch

from helpers import load_golden, rel_l2
from oracle import seeding, vss_ref

CASES = ["c128_12x16", "c256_9x9", "c512_8x10"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_block_matches_reference(name):
    gold = load_golden("vss")["cases"][name]
    c, b, h, w = gold["shape"]
    from tamtr_b200.vss import VSSBlock
    blk = VSSBlock(hidden_dim=c, drop_path=0.0)                     # construction on the CPU is allowed; forward is not
    assert {k: tuple(v.shape) for k, v in blk.state_dict().items()} == gold["manifest"]
    vss_ref.seed_block(blk, gold["param_seed"])
    sd = {"b." + k: v.detach().clone().requires_grad_() for k, v in blk.state_dict().items()}
    x = seeding.seeded_tensor(600 + c, "x", (b, h, w, c)).requires_grad_()
    probe = seeding.seeded_tensor(600 + c, "probe", (b, h, w, c))
    y = vss_ref.vss_block(sd, "b", x)
    (y * probe).sum().backward()
    assert rel_l2(y, gold["y"]) < 1e-5 and rel_l2(x.grad, gold["grad_x"]) < 1e-5
    assert rel_l2(sd["b.op.A_logs"].grad, gold["grad_A_logs"]) < 1e-4
    assert rel_l2(sd["b.op.x_proj_weight"].grad, gold["grad_x_proj"]) < 1e-4
    assert rel_l2(sd["b.op.dt_projs_bias"].grad, gold["grad_dt_bias"]) < 1e-4


def test_scan_restatement_gradcheck():
    torch.manual_seed(0)
    b, k, d, n, l = 1, 2, 2, 3, 5
    u = torch.randn(b, k * d, l, dtype=torch.float64, requires_grad=True)
    delta = torch.randn(b, k * d, l, dtype=torch.float64, requires_grad=True)
    A = (-torch.rand(k * d, n, dtype=torch.float64) - 0.2).requires_grad_()
    B = torch.randn(b, k, n, l, dtype=torch.float64, requires_grad=True)
    C = torch.randn(b, k, n, l, dtype=torch.float64, requires_grad=True)
    D = torch.randn(k * d, dtype=torch.float64, requires_grad=True)
    bias = torch.randn(k * d, dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(lambda *a: vss_ref.selective_scan(*a, True), (u, delta, A, B, C, D, bias), atol=1e-7)


def test_product_scan_rejects_cpu_tensors():
    from tamtr_b200.vss import selective_scan
    z = torch.zeros(1, 512, 8)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        selective_scan(z, z, torch.zeros(512, 16), torch.zeros(1, 4, 16, 8), torch.zeros(1, 4, 16, 8))


def test_recurrence_oracle_agrees_with_the_closed_form():
    """oracle/vss_ref.selective_scan (the loop over positions) against selective_scan_closed_form (cumulative sums + a
    masked contraction, no loop): outputs and all seven gradients in fp64 -- two formulations that share no code."""
    import torch
    from oracle import seeding, vss_ref
    b, k, d, l, n = 2, 4, 3, 37, 16
    mk = lambda name, shape: seeding.seeded_tensor(17, name, shape).double()
    base = [mk("u", (b, k * d, l)), mk("dt", (b, k * d, l)) - 1.5,
            -(0.5 + 8.0 * seeding.seeded_uniform(17, "A", (k * d, n)).double()),
            mk("B", (b, k, n, l)), mk("C", (b, k, n, l)), mk("D", (k * d,)), 0.3 * mk("bias", (k * d,))]
    gout = mk("g", (b, k * d, l))
    res = []
    for fn in (vss_ref.selective_scan, vss_ref.selective_scan_closed_form):
        leaves = [t.clone().requires_grad_() for t in base]
        y = fn(*leaves)
        y.backward(gout)
        res.append([y.detach()] + [t.grad for t in leaves])
    for a, c in zip(*res):
        assert ((a - c).norm() / c.norm()).item() < 1e-10
