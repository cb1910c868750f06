import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
def run(B,P,T,D,s,path,seed=0):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B,P,D,generator=g).to(torch.bfloat16).cuda().requires_grad_(True)
    l = torch.randn(B,T,D,generator=g).to(torch.bfloat16).cuda().requires_grad_(True)
    m = torch.ones(B,T,dtype=torch.bool).cuda()
    out = SPARCLoss(cfg(1.0/P, s), kernel_path=path)(v,l,m)
    out['total_loss'].backward(); torch.cuda.synchronize()
    return v.grad.float().cpu(), l.grad.float().cpu()
for (B,P,T,D) in [(2,33,20,256),(2,33,77,256),(2,50,20,256),(2,48,32,256),(2,33,20,512),(4,33,20,256)]:
    for rep in range(2):
        dv, dl = run(B,P,T,D,3.0,'tc')
        rv, rl = run(B,P,T,D,3.0,'simt')
        nv = torch.isnan(dv); nl = torch.isnan(dl)
        print((B,P,T,D), 'rep',rep,'nan dv', int(nv.sum()), 'rows', nv.any(-1).nonzero()[:6].tolist(), 'cols', nv.any(1).nonzero()[:4].tolist(), 'nan dl', int(nl.sum()),
              'err dv', float((dv-rv)[~nv].norm()/rv.norm()), 'err dl', float((dl-rl)[~nl].norm()/rl.norm()))
