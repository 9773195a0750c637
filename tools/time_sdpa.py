"""Decoder self-attention (16 images x 8 heads x 300 queries x 64): which library SDPA backend is fastest, forward + backward."""
import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

B, H, L, D = 16, 8, 300, 64
dev = "cuda"
qkv = [torch.randn(B, L, H, D, device=dev, dtype=torch.bfloat16).transpose(1, 2).requires_grad_() for _ in range(3)]
mask = torch.rand(L, L, device=dev) > 0.3
go = torch.randn(B, H, L, D, device=dev, dtype=torch.bfloat16)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION),
                 ("flash(no mask)", SDPBackend.FLASH_ATTENTION), ("math", SDPBackend.MATH)):
    for m in (mask, None):
        if "flash" in name and m is not None:
            continue
        try:
            with sdpa_kernel(be):
                def fwd():
                    return F.scaled_dot_product_attention(*qkv, attn_mask=m)

                def fb():
                    o = fwd()
                    o.backward(go)
                with torch.no_grad():
                    tf = timed(fwd)
                tfb = timed(fb)
            print(f"{name:16s} mask={m is not None}: fwd {tf:7.1f} us   fwd+bwd {tfb:7.1f} us (host-inclusive loop)")
        except Exception as e:
            print(f"{name:16s} mask={m is not None}: {type(e).__name__}: {str(e)[:100]}")

print("---- cuDNN with different mask encodings (forward only, device time)")
neg = torch.zeros(L, L, device=dev, dtype=torch.bfloat16).masked_fill(~mask, float("-inf"))
variants = {"bool [L,L]": mask, "bf16 additive [L,L]": neg, "bf16 additive [1,1,L,L]": neg.view(1, 1, L, L),
            "bf16 additive [B,H,L,L]": neg.view(1, 1, L, L).expand(B, H, L, L).contiguous(),
            "fp32 additive [L,L]": neg.float()}
for be_name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
    for vn, m in variants.items():
        try:
            with sdpa_kernel(be), torch.no_grad():
                t = timed(lambda: F.scaled_dot_product_attention(*qkv, attn_mask=m if m.dtype == torch.bool else m.to(qkv[0].dtype) if m.dtype != torch.float32 else m.to(qkv[0].dtype)), 100)
            print(f"{be_name:10s} {vn:26s} fwd {t:7.1f} us")
        except Exception as e:
            print(f"{be_name:10s} {vn:26s} {type(e).__name__}: {str(e)[:90]}")

print("---- fwd + bwd as a CUDA graph (device time per replay)")
for be_name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
    for vn in ("bool [L,L]", "bf16 additive [L,L]"):
        m = variants[vn]
        try:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s), sdpa_kernel(be):
                for _ in range(3):
                    for t in qkv:
                        t.grad = None
                    F.scaled_dot_product_attention(*qkv, attn_mask=m).backward(go)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            for t in qkv:
                t.grad = None
            with sdpa_kernel(be), torch.cuda.graph(g):
                F.scaled_dot_product_attention(*qkv, attn_mask=m).backward(go)
            t = timed(g.replay, 100)
            print(f"{be_name:10s} {vn:22s} fwd+bwd {t:7.1f} us")
        except Exception as e:
            print(f"{be_name:10s} {vn:22s} {type(e).__name__}: {str(e)[:120]}")

print("---- csrc/selfattn.cu on the packed projections (CUDA graph, device time per replay)")
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tamtr_b200 import ops  # noqa: E402
d = H * D
qkp = torch.randn(B, L, 2 * d, device=dev, dtype=torch.bfloat16).requires_grad_()
vp = torch.randn(B, L, d, device=dev, dtype=torch.bfloat16).requires_grad_()
gop = torch.randn(B, L, d, device=dev, dtype=torch.bfloat16)
blocked = ops.attention_mask_bits(~mask)
for what in ("fwd", "fwd+bwd"):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            qkp.grad = vp.grad = None
            o = ops._SelfAttnFn.apply(qkp, vp, blocked, H)
            if what != "fwd":
                o.backward(gop)
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    qkp.grad = vp.grad = None
    with torch.cuda.graph(g):
        o = ops._SelfAttnFn.apply(qkp, vp, blocked, H)
        if what != "fwd":
            o.backward(gop)
    print(f"ours {what:8s} {timed(g.replay, 100):7.1f} us")
