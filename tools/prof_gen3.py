"""clock64 phase stamps of the third-generation SPARC kernels (median over CTAs; device profile buffer).
usage: python tools/prof_gen3.py [B] [cold]"""
import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, _lib
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
cold = len(sys.argv) > 2
P, T, D = 196, 77, 512
torch.manual_seed(0)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P))
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
names = {'fwd': (['start','P0 issued','w_ready seen','P1 issued'],
                 ['start','side job done','norms + pooled done','s_full seen','sweep 1 done','sweep 2 done (w_ready)','sigma / stats done','G blk 0 done','all G + csc done','E3 logits in smem','end']),
         'bwd': (['start','P1 issued','e1_ready seen','ds_ready seen','P4 issued'],
                 ['start','phase 0 done','s_full seen','E1 done','dw_full seen','E3 done','dl blk 0 done','dv blk 0 done','end'])}
for which, setter in (('fwd', _lib.lib.cfa_debug_set_profile_buffer_fwd), ('bwd', _lib.lib.cfa_debug_set_profile_buffer)):
    buf = torch.zeros(B,32,dtype=torch.int64,device='cuda')
    v.grad=None; l.grad=None
    if cold: flush.zero_()
    setter(buf.data_ptr())
    out = crit(v,l,m)['total_loss']
    if cold and which == 'bwd': flush.zero_()
    out.backward()
    torch.cuda.synchronize()
    setter(0)
    t = buf.cpu().double()
    for wname, sel in (('wave 1', slice(0, min(B, 148))), ('wave 2', slice(148, B))):
        if sel.start >= B: continue
        mma = t[sel, :8]; epi = t[sel, 16:30]; t0 = mma[:, 0:1]
        print(f'{which.upper()} B = {B} {"L2-cold" if cold else "warm"} {wname}: MMA thread (cycles since start, median over CTAs)')
        for i, n in enumerate(names[which][0]): print(f'  {n:20s} {float((mma[:, i:i+1] - t0).median()):10.0f}')
        if which == 'bwd': print('  P4 MMA waits: full %.0f  out_free %.0f' % (float(t[sel, 8].median()), float(t[sel, 9].median())))
        print(' epilogue thread 0:')
        for i, n in enumerate(names[which][1]): print(f'  {n:24s} {float((epi[:, i:i+1] - t0).median()):10.0f}')
        if which == 'fwd': print('  E3 probes since start (gn2 sums written, l_full seen, logits loop done, LSE dir 0, LSE dir 1):', [float((t[sel, 16+k:16+k+1] - t0).median()) for k in (11, 12, 13, 14, 15)])
