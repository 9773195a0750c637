import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_step_gpu import _setup, _loss, _grads
from tamtr_b200 import dp, ops
m, xs, text, plan = _setup()
sd = {k: v.clone() for k, v in m.state_dict().items()}
xc = [x.cuda() for x in xs]
for it in range(3):
    m.load_state_dict(sd)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        feats, shapes, hub = m._encode(xc)
        anchors, valid = m._anchors(shapes, feats.dtype, feats.device)
        rank = m._rank_tokens(feats, valid)
        with torch.no_grad():
            ref = m.enc_score_head(m.enc_output(valid * feats)).max(-1).values.float()
    torch.cuda.synchronize()
    d = (rank - ref).abs()
    print(it, 'rank vs torch: max diff', d.max().item(), 'n bad', (d > 0.05).sum().item(), 'valid sum', valid.sum().item(),
          'u8 sum', m._anchor_cache[("valid_u8", valid.data_ptr())].sum().item(), 'ref max', ref.max().item(), 'rank max', rank.max().item())
    bad = (d > 0.05).nonzero()
    if len(bad): print('   first bad', bad[:5].tolist(), 'valid there', valid.view(-1)[bad[:5, 1]].tolist())
