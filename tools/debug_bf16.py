import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import *
from oracle import head_ref, seeding
from tamtr_b200.head import ManbaWorldDecoder
torch.backends.cudnn.allow_tf32 = False
c = load_golden("modules_heads")["cases"]["meh_syaml_small"]
m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False)
filled_state_dict(m, 73, c["manifest"])
m.cuda().train()
sd = {k: v.detach() for k, v in m.state_dict().items()}
for B, Lq, shapes in ((4, 120, [[40, 40], [20, 20], [10, 10]]), (4, 300, [[160, 160], [80, 80], [40, 40]])):
    Lv = sum(h * w for h, w in shapes)
    embed = seeding.seeded_tensor(81, "embed", (B, Lq, 512)).cuda()
    feats = seeding.seeded_tensor(81, "feats", (B, Lv, 512)).cuda()
    refer = torch.logit(torch.cat([seeding.seeded_uniform(81, "xy", (B, Lq, 2), 0.05, 0.95), seeding.seeded_uniform(81, "wh", (B, Lq, 2), 0.02, 0.3)], -1)).cuda()
    text = torch.nn.functional.normalize(seeding.seeded_tensor(81, "text", (B, 10, 512)), dim=-1).cuda()
    res = {}
    for impl in ("product", "torch_ops"):
        for mode in ("fp32", "bf16"):
            e, f = embed.clone().requires_grad_(), feats.clone().requires_grad_()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                if impl == "product":
                    db, ds = m.decoder(e, refer, f, shapes, text, m.dec_bbox_head, m.dec_score_head, m.query_pos_head)
                else:
                    db, ds = head_ref.decoder(sd, "", e, refer, f, shapes, 3, 8, True, text=text)
            (probe_loss(db.float(), 82, "pb") + 0.01 * probe_loss(ds.float(), 82, "ps")).backward()
            res[impl, mode] = (db.detach().float(), ds.detach().float(), e.grad.float(), f.grad.float())
    print(shapes[0], 'product bf16 vs product fp32   ', ['%.4f' % rel_l2(a, b) for a, b in zip(res["product", "bf16"], res["product", "fp32"])])
    print(shapes[0], 'torch-ops bf16 vs torch-ops fp32', ['%.4f' % rel_l2(a, b) for a, b in zip(res["torch_ops", "bf16"], res["torch_ops", "fp32"])])
    print(shapes[0], 'product fp32 vs torch-ops fp32  ', ['%.2e' % rel_l2(a, b) for a, b in zip(res["product", "fp32"], res["torch_ops", "fp32"])])
    print(shapes[0], 'product bf16 vs torch-ops bf16  ', ['%.4f' % rel_l2(a, b) for a, b in zip(res["product", "bf16"], res["torch_ops", "bf16"])])
