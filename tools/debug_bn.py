import torch, torch.nn as nn
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(0)
class CG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x): return x.view_as(x)
    @staticmethod
    def backward(ctx, g):
        print('   grad strides in', g.shape, g.stride(), g.is_contiguous())
        return g.contiguous()
B = 1
conv = nn.Conv2d(64, 128, 1, bias=False); bn = nn.BatchNorm2d(128)
x = torch.randn(B, 64, 20, 20)
g = torch.randn(B, 400, 128)
def run(dev, variant):
    c, b = conv.to(dev), bn.to(dev)
    c.zero_grad(); b.zero_grad()
    xi = x.detach().clone().to(dev).requires_grad_()
    y = c(xi)
    y.retain_grad()
    f = b(y)
    f.retain_grad()
    if variant == 'cg': f2 = CG.apply(f)
    elif variant == 'clone': f2 = f.clone()
    else: f2 = f
    out = f2.flatten(2).permute(0, 2, 1)
    if variant == 'mulsum':
        (out * g.to(dev)).sum().backward()
    else:
        out.backward(g.to(dev))
    return xi.grad.cpu(), b.bias.grad.cpu().clone(), b.weight.grad.cpu().clone(), f.grad.cpu(), y.grad.cpu(), c.weight.grad.cpu().clone()
for variant in ('plain', 'cg', 'clone', 'mulsum'):
    ref = run('cpu', variant); got = run('cuda', variant)
    print(variant, ['%.2e' % ((a-b).norm()/b.norm()).item() for a, b in zip(got, ref)])
