import torch
dev='cuda'
B,C,HW,d=16,128,25600,512
x=torch.randn(B,C,HW,device=dev).bfloat16()
dpre=torch.randn(B,33600,d,device=dev).bfloat16(); dp=dpre[:, :HW]
def t(fn,n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/n*1e3
def f_bmm(): return torch.bmm(x, dp).float().sum(0)
def f_einsum(): return torch.einsum('bch,bhd->cd', x, dp)
def f_loop():
    out=torch.zeros(C,d,device=dev,dtype=torch.float32)
    for b in range(B): out += torch.mm(x[b], dp[b], out_dtype=torch.float32)
    return out
def f_loop2():
    out=torch.mm(x[0], dp[0])
    for b in range(1,B): out.addmm_(x[b], dp[b])
    return out
def f_cat():   # one GEMM over K = B*HW: needs x as [C, B*HW]
    xt = x.transpose(0,1).reshape(C, B*HW)
    return torch.mm(xt, dp.reshape(B*HW, d))
ref=f_bmm()
for f in (f_einsum, f_loop, f_loop2, f_cat): print(f.__name__, ((f().float()-ref).norm()/ref.norm()).item())
print('bmm+sum %.1f  einsum %.1f  loop mm fp32 %.1f  loop addmm_ %.1f  cat-K %.1f us' % (t(f_bmm), t(f_einsum), t(f_loop), t(f_loop2), t(f_cat)))
