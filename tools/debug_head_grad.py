import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import *
from oracle import head_ref, seeding
from oracle.make_goldens import _synthetic_targets
from tamtr_b200.head import ManbaWorldDecoder, get_cdn_group
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
c = load_golden("modules_heads")["cases"]["meh_syaml_full"]
for sizes, B in (((160, 80, 40), 1), ((40, 20, 10), 2), ((80, 40, 20), 1)):
    m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False)
    filled_state_dict(m, 73, None)
    sd = {k: v.detach().clone().cpu() for k, v in m.state_dict().items()}
    sd = {k: (v.requires_grad_() if v.is_floating_point() and "running" not in k else v) for k, v in sd.items()}
    m.cuda().train()
    xs_c = [seeding.seeded_smooth_map(74, f"x{i}", (B, ch, s, s)).requires_grad_() for i, (ch, s) in enumerate(zip((128, 256, 512), sizes))]
    xs = [x.detach().cuda().requires_grad_() for x in xs_c]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(74, "text", (B, 10, 512)), dim=-1)
    batch = _synthetic_targets(75, B, 5, 20)
    torch.manual_seed(1234)
    db, ds, eb, es, meta = m(xs, text.cuda(), batch)
    feats_hook = {}
    loss = head_ref.surrogate_loss(db, ds, eb, es); loss.backward()
    torch.manual_seed(1234)
    dn_embed, dn_bbox, attn_mask, _ = get_cdn_group(batch, 10, 100, sd["denoising_class_embed.weight"], 100, 0.5, 1.0, True)
    rdb, rds, reb, res = head_ref.head(sd, "", xs_c, 100, 3, 8, training=True, text=text, cdn=(dn_embed, dn_bbox, attn_mask))
    rl = head_ref.surrogate_loss(rdb, rds, reb, res); rl.backward()
    print(sizes, B, 'fwd', rel_l2(db, rdb), rel_l2(ds, rds), rel_l2(eb, reb), 'loss', loss.item(), rl.item())
    for i in range(3):
        print('  grad x%d' % i, rel_l2(xs[i].grad, xs_c[i].grad), xs[i].grad.norm().item(), xs_c[i].grad.norm().item())
    worst = sorted(((rel_l2(p.grad, sd[k].grad), k) for k, p in m.named_parameters() if p.grad is not None and sd[k].grad is not None and sd[k].grad.norm() > 0), reverse=True)[:8]
    for e, k in worst:
        print('  param', k, e)
