"""Per-CTA durations and phase medians of sparc_bwd2_kernel at the bench shape (B = 256: two waves), inputs L2-cold
(a 512 MB buffer is written before the profiled step), split by wave."""
import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, _lib
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B,P,T,D = (int(sys.argv[1]) if len(sys.argv) > 1 else 256),196,77,512
torch.manual_seed(0)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P))
flush = torch.empty(512 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
for which, setter in (('bwd', _lib.lib.cfa_debug_set_profile_buffer), ('fwd', _lib.lib.cfa_debug_set_profile_buffer_fwd)):
    buf = torch.zeros(B,32,dtype=torch.int64,device='cuda')
    v.grad=None; l.grad=None
    flush.zero_()
    setter(buf.data_ptr())
    out = crit(v,l,m)['total_loss']
    if which == 'bwd':
        flush.zero_()
    out.backward()
    torch.cuda.synchronize()
    setter(0)
    t = buf.cpu().double()
    mma = t[:, :8]; epi = t[:, 16:30]
    start = mma[:, 0]
    last = epi.max(dim=1).values
    dur = last - start
    g0 = start.min()
    print(f'{which}: B = {B}')
    for name, sel in (('wave 1 (b < 148)', slice(0, 148)), ('wave 2', slice(148, B))):
        if sel.start >= B: continue
        d = dur[sel]; s0 = start[sel] - g0; e0 = last[sel] - g0
        print(f'  {name:18s} CTA duration min/median/max {d.min():9.0f} {d.median():9.0f} {d.max():9.0f} | start median {s0.median():9.0f} max {s0.max():9.0f} | end median {e0.median():9.0f} max {e0.max():9.0f}')
        ph = (mma[sel] - start[sel, None]).median(dim=0).values
        print('     MMA-thread stamps (median):', [int(x) for x in ph[:7]])
        pe = (epi[sel] - start[sel, None]).median(dim=0).values
        print('     epilogue stamps   (median):', [int(x) for x in pe[:12]])
