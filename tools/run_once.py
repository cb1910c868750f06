"""One warm step of the SPARC loss at the bench shape (for ncu captures): python tools/run_once.py [B]"""
import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
P, T, D = 196, 77, 512
torch.manual_seed(0)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(types.SimpleNamespace(similarity_threshold=1.0/P, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=1.0))
for _ in range(3):
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
torch.cuda.synchronize()
