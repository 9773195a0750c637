import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import *
from oracle import head_ref, seeding
from tamtr_b200.head import ManbaWorldDecoder
torch.backends.cudnn.allow_tf32 = False
m = ManbaWorldDecoder(10, [128, 256, 512], 512, 100, 4, 8, 3, vss=False)
filled_state_dict(m, 73, None)
m.cuda().train(); m.num_denoising = 0
sizes = (40, 20, 10)
res = {}
for B in (1, 2):
    xs = [seeding.seeded_tensor(74, f"x{i}", (1, ch, s, s)).repeat(B, 1, 1, 1).cuda().requires_grad_() for i, (ch, s) in enumerate(zip((128, 256, 512), sizes))]
    text = torch.nn.functional.normalize(seeding.seeded_tensor(74, "text", (1, 10, 512)), dim=-1).repeat(B, 1, 1).cuda()
    cap = {}
    orig = m._get_encoder_input
    def wrapped(x):
        feats, shapes = orig(x)
        feats.register_hook(lambda g: cap.__setitem__('gfeats', g.detach().clone()))
        cap['feats'] = feats.detach().clone()
        return feats, shapes
    m._get_encoder_input = wrapped
    m.zero_grad()
    db, ds, eb, es, meta = m(xs, text, None)   # no CDN -> deterministic
    loss = head_ref.surrogate_loss(db, ds, eb, es); loss.backward()
    m._get_encoder_input = orig
    res[B] = dict(loss=loss.item(), gx=[x.grad[0].clone() for x in xs], gfeats=cap['gfeats'][0], feats=cap['feats'][0],
                  gp={k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
print('loss', res[1]['loss'], res[2]['loss'])
print('feats', rel_l2(res[1]['feats'], res[2]['feats']))
print('gfeats (B1 vs 2*B2[0])', rel_l2(res[1]['gfeats'], 2 * res[2]['gfeats']))
for i in range(3):
    print('gx', i, rel_l2(res[1]['gx'][i], 2 * res[2]['gx'][i]))
worst = sorted(((rel_l2(res[1]['gp'][k], res[2]['gp'][k]), k) for k in res[1]['gp'] if res[2]['gp'][k].norm() > 0), reverse=True)[:6]
print(worst)
