import sys, types, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss
from oracle import losses_oracle as lo
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
def rel(x, r): return float((x.double().cpu() - r).norm() / r.norm())
for (B, P, T, D, s) in [(3, 196, 77, 512, 1.0), (2, 197, 77, 512, 2.0), (4, 50, 77, 256, 1.0), (2, 33, 20, 256, 3.0), (5, 100, 64, 256, 1.0)]:
    g = torch.Generator().manual_seed(B * 1000 + P)
    v0 = torch.randn(B, P, D, generator=g).to(torch.bfloat16); l0 = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool)
    if B == 4: m[1, 60:] = False; m[3, 5:] = False
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    v = v0.cuda().requires_grad_(True); l = l0.cuda().requires_grad_(True)
    out = SPARCLoss(cfg(1.0 / P, s), kernel_path='tc')(v, l, m.cuda())
    out['total_loss'].backward(); torch.cuda.synchronize()
    o = lo.sparc_forward(v0.double(), l0.double(), m, thr, 1.0, 1.0, s, mask_semantics='truncate')
    rv, rl = lo.sparc_backward(o)
    print((B, P, T, D, s), 'loss', float(out['total_loss']), float(o['total_loss']), 'dv', rel(v.grad.float(), rv), 'dl', rel(l.grad.float(), rl),
          'nan', int(torch.isnan(v.grad).sum()), int(torch.isnan(l.grad).sum()), flush=True)
