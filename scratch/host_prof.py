import sys, types, time, torch, cProfile, pstats
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, _lib
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B,P,T,D = 256,196,77,512
torch.manual_seed(0)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P))
def step():
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
for _ in range(5): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(22)
