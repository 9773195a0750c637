"""TEST INFRASTRUCTURE ONLY -- deterministic, name-keyed parameter fill.

Module-level goldens would need the reference's multi-MB state_dicts; instead both sides (the reference module
in oracle/make_goldens.py and the product module in tests/) get their parameters/buffers overwritten by this
function, which depends only on the state_dict KEY, the tensor SHAPE and a seed.  That the keys and shapes agree
is exactly the drop-in contract (checkpoints of the reference must load, SURVEY.md section 5 "Checkpoint"), and the
fixture stores the key->shape manifest so the tests assert it.
"""
import zlib

import torch


def _gen(seed, name):
    return torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 31))


@torch.no_grad()
def seeded_fill(module, seed):
    """Overwrite every floating-point entry of module.state_dict() in place; returns {key: shape}."""
    manifest = {}
    for name, t in module.state_dict().items():
        manifest[name] = tuple(t.shape)
        if not t.is_floating_point():
            continue
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        shape = tuple(t.shape)
        if leaf == "running_var":
            new = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            new = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "logit_scale" or (leaf == "bias" and t.numel() == 1 and "score_head" in name):
            continue                                        # ContrastiveHeadMLP scalars keep their init constants
        elif name.endswith("sampling_offsets.bias"):
            new = t.detach().cpu().float() + 0.3 * torch.randn(shape, generator=g)   # ring grid + jitter
        elif name.endswith("sampling_offsets.weight") or name.endswith("attention_weights.weight"):
            new = 0.05 * torch.randn(shape, generator=g)    # zero at init (transformer.py:236,245) -> query dependent
        elif t.dim() >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            new = torch.randn(shape, generator=g) / max(1.0, fan_in) ** 0.5
        elif leaf == "weight":                              # norm scales
            new = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                               # biases and other vectors
            new = 0.1 * torch.randn(shape, generator=g)
        t.copy_(new.to(t.dtype))
    return manifest


def seeded_tensor(seed, name, shape, scale=1.0):
    return scale * torch.randn(tuple(shape), generator=_gen(seed, name))


def seeded_uniform(seed, name, shape, lo=0.0, hi=1.0):
    return lo + (hi - lo) * torch.rand(tuple(shape), generator=_gen(seed, name))


def seeded_smooth_map(seed, name, shape, factor=4):
    """Smooth synthetic feature map [B,C,H,W]: a coarse seeded random grid upsampled bilinearly.  Real pyramid
    features are spatially smooth; with white noise the derivative of a bilinear sample w.r.t. its location is pure
    noise of size W * |v|, which makes every gradient that flows through the sampling offsets chaotic in fp32 (the
    reference's own fp32 and fp64 runs then disagree at the percent level) and useless as a parity target."""
    B, C, H, W = shape
    coarse = torch.randn(B, C, max(2, H // factor), max(2, W // factor), generator=_gen(seed, name))
    return torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=True).contiguous()


def seeded_smooth_tokens(seed, name, B, d, shapes, factor=4):
    """Token tensor [B, sum(h*w), d] whose levels are smooth maps (seeded_smooth_map), flattened row-major and
    concatenated in pyramid order like head.py:1210-1218."""
    levels = [seeded_smooth_map(seed, f"{name}{i}", (B, d, int(h), int(w)), factor) for i, (h, w) in enumerate(shapes)]
    return torch.cat([m.flatten(2).permute(0, 2, 1) for m in levels], 1).contiguous()
