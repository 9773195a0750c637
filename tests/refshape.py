"""Reference-SHAPED stand-ins for the classes patch.enable() rebinds (the reference itself does not travel to the GPU
box): built attribute for attribute like the reference's constructors -- ultralytics/nn/modules/head.py:1081-1128,
nn/modules/transformer.py:838-848, nn/extra_modules/VManba/vmamba.py:330-470, 1169-1234 -- with NONE of our helper
methods, so that binding our functions onto them the way enable() does proves those functions only rely on what the
reference provides.  Used by tests/test_patch_gpu.py and by bench.py's `through_enable` field."""
import torch.nn as nn


class RefSS2D(nn.Module):
    """Attributes of SS2D.__initv2__ for forward_type "v2" (extra_modules/VManba/vmamba.py:330-470); like the reference,
    the instance binds its own `forward` and `forward_core` at construction -- and cannot run them (missing extension)."""

    def __init__(self, ours):
        super().__init__()
        import functools
        self.d_conv, self.channel_first = 3, False
        self.disable_force32 = self.oact = self.disable_z = self.disable_z_act = False
        self.out_norm_shape = "v0"
        for name in ("in_proj", "act", "conv2d", "out_norm", "out_proj", "dropout"):
            setattr(self, name, getattr(ours, name))
        for name in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
            setattr(self, name, nn.Parameter(getattr(ours, name).detach().clone()))
        self.out_act = nn.Identity()
        self.forward = self.forwardv2
        self.forward_core = functools.partial(self.forward_corev2, force_fp32=True, SelectiveScan=None)

    def forward_corev2(self, x, **kwargs):
        raise NameError("name 'selective_scan_cuda_core' is not defined")      # what the reference does here

    def forwardv2(self, x, **kwargs):
        return self.forward_core(x)


class RefVSSBlock(nn.Module):
    def __init__(self, ours):
        super().__init__()
        self.ssm_branch = self.mlp_branch = True
        self.use_checkpoint = self.post_norm = False
        self.norm, self.op, self.drop_path = ours.norm, RefSS2D(ours.op), ours.drop_path
        self.norm2, self.mlp = ours.norm2, ours.mlp

    def forward(self, input):
        return input + self.drop_path(self.op(self.norm(input)))


class RefTextDecoder(nn.Module):               # transformer.py:838-848
    def __init__(self, hidden_dim, layers, eval_idx=-1):
        super().__init__()
        self.layers = layers
        self.num_layers = len(layers)
        self.hidden_dim = hidden_dim
        self.eval_idx = eval_idx if eval_idx >= 0 else self.num_layers + eval_idx


class RefMEH(nn.Module):                       # nn/modules/head.py:1081-1128, attribute for attribute
    export = False

    def __init__(self, ours):
        super().__init__()
        self.hidden_dim, self.nhead, self.nl, self.nc = ours.hidden_dim, ours.nhead, ours.nl, ours.nc
        self.num_queries, self.num_decoder_layers = ours.num_queries, ours.num_decoder_layers
        self.input_proj = ours.input_proj
        self.VSSBlocks = nn.ModuleList(b if isinstance(b, nn.Identity) else RefVSSBlock(b) for b in ours.VSSBlocks)
        if any(isinstance(b, nn.Identity) for b in ours.VSSBlocks):
            self.vss = False                    # (our `vss=False` test configuration; the reference has no such switch)
        self.num_Blocks = len(self.VSSBlocks)
        self.decoder = RefTextDecoder(ours.hidden_dim, ours.decoder.layers)
        self.denoising_class_embed = ours.denoising_class_embed
        self.num_denoising, self.label_noise_ratio = ours.num_denoising, ours.label_noise_ratio
        self.box_noise_scale, self.learnt_init_query = ours.box_noise_scale, ours.learnt_init_query
        self.query_pos_head, self.enc_output = ours.query_pos_head, ours.enc_output
        self.enc_score_head, self.enc_bbox_head = ours.enc_score_head, ours.enc_bbox_head
        self.dec_score_head, self.dec_bbox_head = ours.dec_score_head, ours.dec_bbox_head


_installed = False


def install_like_enable():
    """The rebinding patch.enable() performs on the reference's classes, applied to the stand-ins (once)."""
    global _installed
    if _installed:
        return
    from tamtr_b200 import head, modules, patch, vss
    patch.install(RefMEH, head.ManbaWorldDecoder, patch.HEAD_ATTRS)
    patch.install(RefTextDecoder, modules.TextDeformableTransformerDecoder, patch.DECODER_ATTRS)
    RefVSSBlock.forward = vss.vssblock_forward_on(RefVSSBlock.forward)
    RefSS2D.forwardv2 = vss.ss2d_forward_on(RefSS2D.forwardv2)
    _installed = True
