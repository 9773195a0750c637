"""Experiment: how much of the sampler's time is the value LAYOUT?  Same gathers, three layouts of `value`:
  token-major   [B, Lv, H, Dh]            (the reference's layout after value_proj; a (q, h) reads 128 B every 1 KB)
  arena slice   [B, Lv, 3*H*Dh] column slice (batched projection of 3 layers: 128 B every 3 KB)
  head-major    [B*H, Lv, 1, Dh]          (each head its own slab: neighbouring pixels of a head are contiguous)
The head-major run is the SAME kernel called with H = 1 and B*H "images"."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tamtr_b200  # noqa: E402
from tamtr_b200 import _lib  # noqa: E402


def time_fn(fn, iters, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for s, e in ev:
        flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    B, Lq, H, Dh, base = 16, 300, 8, 64, 160
    shapes = [[base, base], [base // 2, base // 2], [base // 4, base // 4]]
    Lv = sum(h * w for h, w in shapes)
    g = torch.Generator(device=dev).manual_seed(0)
    value = torch.randn(B, Lv, H, Dh, device=dev, generator=g).bfloat16()
    cxcy = torch.rand(B, Lq, 1, 1, 1, 2, device=dev, generator=g)
    wh = 0.01 + 0.29 * torch.rand(B, Lq, 1, 1, 1, 2, device=dev, generator=g)
    loc = (cxcy + torch.randn(B, Lq, H, 3, 4, 2, device=dev, generator=g) * 2.0 / 4 * wh * 0.5).contiguous()
    attn = torch.softmax(torch.randn(B, Lq, H, 12, device=dev, generator=g), -1).view(B, Lq, H, 3, 4)
    gout = torch.randn(B, Lq, H * Dh, device=dev, generator=g).bfloat16()
    sh, _ = _lib.shapes_array(shapes)
    lib = _lib.lib()
    st = _lib.stream_ptr(dev)
    res = {}

    def run(name, v, l, a, go, b, h, tok_stride):
        out = torch.empty(b, Lq, h * Dh, dtype=torch.bfloat16, device=dev)
        gv = torch.zeros(v.shape if tok_stride == 0 else (b, Lv, tok_stride), dtype=torch.bfloat16, device=dev)
        gl, ga = torch.empty_like(l), torch.empty_like(a)
        f = lambda: lib.tamtr_msda_forward(v.data_ptr(), l.data_ptr(), a.data_ptr(), out.data_ptr(), 1, b, Lv, h, Dh, Lq, 3, 4,
                                           sh, tok_stride, st)
        bw = lambda: lib.tamtr_msda_backward(go.data_ptr(), v.data_ptr(), l.data_ptr(), a.data_ptr(), gv.data_ptr(),
                                             gl.data_ptr(), ga.data_ptr(), 1, b, Lv, h, Dh, Lq, 3, 4, sh, tok_stride, 0, None, 1, st)
        assert f() == 0 and bw() == 0
        res[name] = {"fwd_us": time_fn(f, 20, flush), "bwd_us_no_memset": time_fn(bw, 20, flush)}
        return out

    o_tok = run("token_major", value, loc, attn, gout, B, H, 0)
    wide = torch.randn(B, Lv, 3 * H * Dh, device=dev, generator=g).bfloat16()
    wide[:, :, H * Dh:2 * H * Dh] = value.view(B, Lv, H * Dh)
    o_ar = run("arena_slice", wide[:, :, H * Dh:2 * H * Dh], loc, attn, gout, B, H, 3 * H * Dh)
    v_hm = value.permute(0, 2, 1, 3).reshape(B * H, Lv, 1, Dh).contiguous()
    l_hm = loc.permute(0, 2, 1, 3, 4, 5).reshape(B * H, Lq, 1, 3, 4, 2).contiguous()
    a_hm = attn.permute(0, 2, 1, 3, 4).reshape(B * H, Lq, 1, 3, 4).contiguous()
    g_hm = gout.view(B, Lq, H, Dh).permute(0, 2, 1, 3).reshape(B * H, Lq, Dh).contiguous()
    o_hm = run("head_major", v_hm, l_hm, a_hm, g_hm, B * H, 1, 0)
    same = torch.equal(o_tok, o_ar) and torch.equal(o_tok.view(B, Lq, H, Dh), o_hm.view(B, H, Lq, Dh).permute(0, 2, 1, 3))
    res["outputs_identical"] = bool(same)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
