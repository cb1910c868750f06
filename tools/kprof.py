import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss
from torch.profiler import profile, ProfilerActivity
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B,P,T,D = 256,196,77,512
torch.manual_seed(0)
vs = [torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True) for _ in range(4)]
ls = [torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True) for _ in range(4)]
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P))
def step(i):
    v, l = vs[i % 4], ls[i % 4]
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
for i in range(5): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(20): step(i)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(1, e.count), e.count) for e in prof.key_averages() if e.device_time_total > 0]
for k, t, c in sorted(rows, key=lambda r: -r[1] * r[2])[:16]:
    print(f'{k[:80]:80s} {t:9.1f} us x {c}')
