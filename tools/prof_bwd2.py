import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, _lib
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B,P,T,D = (int(sys.argv[1]) if len(sys.argv) > 1 else 148),196,77,512
torch.manual_seed(0)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P))
for _ in range(3):
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
buf = torch.zeros(B,32,dtype=torch.int64,device='cuda')
_lib.lib.cfa_debug_set_profile_buffer(buf.data_ptr())
v.grad=None; l.grad=None
crit(v,l,m)['total_loss'].backward()
torch.cuda.synchronize()
_lib.lib.cfa_debug_set_profile_buffer(0)
t = buf.cpu().double()
mma = t[:,:8]; epi = t[:,16:28]
t0 = mma[:,0:1]
names_m = ['start','P1 issued','e1/dl ready seen','ds_ready seen','P4a issued','P4b issued']
names_e = ['start','phase0 done','s_full seen','E1 done','dw_full seen','E3 sweep done','E3 fixup done','bar','E3 done','dl written','dv written']
print('B =', B, ' MMA thread (cycles since start, median over CTAs):')
for i,n in enumerate(names_m): print(f'  {n:18s} {float(((mma[:,i:i+1]-t0)).median()):10.0f}')
print('  MMA waits: P4a full %.0f free %.0f | P4b full %.0f free %.0f' % tuple(float(t[:,k].median()) for k in (8,9,10,11)))
print('epilogue thread row 0, half 0:')
for i,n in enumerate(names_e): print(f'  {n:18s} {float(((epi[:,i:i+1]-t0)).median()):10.0f}')
