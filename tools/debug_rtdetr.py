import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import *
from oracle import seeding
from tamtr_b200.head import RTDETRDecoder
from tamtr_b200 import modules
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
c = load_golden("modules_heads")["cases"]["rtdetr_eval_sbase"]
m = RTDETRDecoder(nc=10, ch=(256, 256, 256)).eval()
filled_state_dict(m, 71, c["manifest"])
m.cuda()
xs = [seeding.seeded_tensor(72, f"x{i}", (2, 256, s, s)).cuda() for i, s in enumerate((80, 40, 20))]
res = {}
for name, (fi, bv, sq) in {"all_off": (False, False, False), "fused_input": (True, False, False), "batched_value": (False, True, False),
                           "sparse_sel": (False, False, True), "all_on": (True, True, True)}.items():
    m.fused_input_proj = fi; m.decoder.batched_value_projection = bv; m.sparse_query_selection = sq
    with torch.no_grad():
        y, (db, ds, eb, es, _) = m(xs)
    res[name] = (db, ds, eb, es)
    print(name, 'vs golden: db %.2e ds %.2e eb %.2e es %.2e' % (rel_l2(db, c["dec_bboxes"]), rel_l2(ds, c["dec_scores"]), rel_l2(eb, c["enc_bboxes"]), rel_l2(es, c["enc_scores"])))
for name in res:
    print(name, 'vs all_off: db %.2e eb %.2e' % (rel_l2(res[name][0], res["all_off"][0]), rel_l2(res[name][2], res["all_off"][2])))
print("---- rank_tokens vs torch")
from tamtr_b200 import ops
m.fused_input_proj = True; m.sparse_query_selection = True
with torch.no_grad():
    feats, shapes = m._get_encoder_input(xs)
    anchors, valid = m._anchors(shapes, feats.dtype, feats.device)
    ref = m.enc_score_head(m.enc_output(valid * feats)).max(-1).values
    mine = m._rank_tokens(feats, valid)
    print('max abs diff', (ref - mine).abs().max().item(), 'ref range', ref.min().item(), ref.max().item())
    bad = (ref - mine).abs() > 1e-3
    print('bad count', bad.sum().item(), 'of', bad.numel(), 'invalid tokens', (~valid.view(-1)).sum().item())
    if bad.any():
        idx = bad.nonzero()[:10]
        print(idx.tolist(), ref[bad][:10].tolist(), mine[bad][:10].tolist(), valid.view(-1)[idx[:, 1]].tolist())
    t1 = torch.topk(ref, 300, dim=1).indices; t2 = torch.topk(mine, 300, dim=1).indices
    print('topk set diff', [len(set(a.tolist()) ^ set(b.tolist())) for a, b in zip(t1, t2)], 'order equal', (t1 == t2).float().mean().item())
