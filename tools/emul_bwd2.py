"""CPU emulation of the restructured tensor-core backward (sparc_bwd2): same phases / operands as the kernel,
checked against the oracle.  split='none' (exact), 'hilo' (bf16 hi+lo operands), 'fp16', 'bf16'."""
import sys, torch
sys.path.insert(0, '.')
from oracle import losses_oracle as lo

def rnd(x, split):
    if split == 'none': return x
    if split == 'hilo':
        hi = x.to(torch.bfloat16).to(x.dtype); lo_ = (x - hi).to(torch.bfloat16).to(x.dtype); return hi + lo_
    if split == 'fp16': return x.to(torch.float16).to(x.dtype)
    if split == 'bf16': return x.to(torch.bfloat16).to(x.dtype)
    if split in ('h16hilo', 'h16hilo_scaled'):
        # fp16 hi + lo operands (what a same-format kind::f16 MMA needs when the raw embeddings are fp16).  'scaled':
        # one power-of-two factor per sample puts the tile's maximum at 2^10 before the split (exact), undone afterwards.
        if split == 'h16hilo_scaled':
            amax = x.abs().flatten(1).max(dim=1).values.clamp_min(1e-300)
            k = torch.floor(10.0 - torch.log2(amax)).view(-1, *([1] * (x.dim() - 1)))
            sc = torch.pow(torch.tensor(2.0, dtype=x.dtype), k)
        else:
            sc = torch.ones((), dtype=x.dtype)
        y = x * sc
        hi = y.to(torch.float16).to(x.dtype); lo_ = (y - hi).to(torch.float16).to(x.dtype)
        return (hi + lo_) / sc

def emul(v, l, mask, thr, s, c_r, c_c, dpool_v, dpool_l, split='none', dt=torch.float64):
    B, P, D = v.shape; T = l.shape[1]
    v = v.to(dt); l = l.to(dt); mf = mask.to(dt)
    il = 1.0 / l.norm(dim=-1).clamp_min(1e-12); ivn = 1.0 / v.norm(dim=-1).clamp_min(1e-12)
    # ---- forward quantities the kernel gets from the forward kernel
    S_raw = torch.einsum('btd,bpd->btp', l, v)
    sn = S_raw * il[:, :, None] * ivn[:, None, :]
    mn, imn = sn.min(-1); mx, imx = sn.max(-1)
    rng = mx - mn + 1e-8
    nn = (sn - mn[..., None]) / rng[..., None]
    kept = ~(nn < thr)
    sigma = torch.where(kept, nn, torch.zeros_like(nn)).sum(-1).clamp_min(1e-8)
    W = torch.where(kept, nn, torch.zeros_like(nn)) / sigma[..., None] * mf[..., None]
    Wq = rnd(W, split)
    G = torch.einsum('btp,bpd->btd', Wq, v) * mf[..., None]
    Gq = rnd(G, split)                                   # what the forward stores (hi/lo)
    ig = 1.0 / G.norm(dim=-1).clamp_min(1e-12)
    y = torch.einsum('bid,bjd->bij', Gq, l) * s * ig[:, :, None] * il[:, None, :]
    m2 = mask[:, :, None] & mask[:, None, :]
    ym = y.masked_fill(~m2, -float('inf'))
    lse_r = torch.logsumexp(ym, 2); lse_c = torch.logsumexp(ym, 1)
    # ---- backward
    g = c_r * torch.exp(ym - lse_r[:, :, None]) + c_c * torch.exp(ym - lse_c[:, None, :])
    g = torch.where(m2, g, torch.zeros_like(g))
    ar = torch.arange(T)
    g[:, ar, ar] -= (c_r + c_c) * mf
    y0 = torch.where(m2, y, torch.zeros_like(y))
    gdot = (g * y0).sum(2); ldot = (g * y0).sum(1)
    gfac = gdot * ig * ig
    dLh = rnd(s * g * ig[:, :, None] * il[:, None, :], split)
    Q = torch.einsum('btd,bpd->btp', Gq, v)
    dWa = torch.einsum('bij,bjp->bip', dLh, rnd(S_raw, split))
    dW = dWa - gfac[..., None] * Q
    Z = torch.einsum('bij,bip->bjp', dLh, Wq)
    wdot = (W * dW).sum(-1)            # kernel re-reads W = hi + lo
    dn = torch.where(kept, (dW - wdot[..., None]) / sigma[..., None], torch.zeros_like(dW))
    nnk = W * sigma[..., None]
    a1 = (dn * (nnk - 1.0)).sum(-1); a2 = (dn * nnk).sum(-1)
    dmn = a1 / rng; dmx = -a2 / rng
    ds = dn / rng[..., None]
    ds.scatter_add_(2, imn[..., None], dmn[..., None]); ds.scatter_add_(2, imx[..., None], dmx[..., None])
    ds = ds * mf[..., None]
    sv = torch.where(kept, nnk * rng[..., None] + mn[..., None], mn[..., None].expand_as(nnk))
    prod = ds * sv
    sdot = prod.sum(-1); vdot = prod.sum(1)
    dSp = rnd(ds * il[:, :, None] * ivn[:, None, :] + Z, split)
    lfac = (sdot + ldot) * il * il; vfac = vdot * ivn * ivn
    Wg = rnd(-gfac[..., None] * W, split)
    cnt = mask.sum(-1, keepdim=True).clamp(min=1e-8).to(dt)
    dv = torch.einsum('btp,btd->bpd', dSp, l) + torch.einsum('btp,btd->bpd', Wg, Gq) - v * vfac[..., None] + dpool_v[:, None, :] / P
    dl = torch.einsum('btp,bpd->btd', dSp, v) - l * lfac[..., None] + mf[..., None] * dpool_l[:, None, :] / cnt[:, :, None]
    return dv, dl

if __name__ == '__main__':
    torch.manual_seed(0)
    for (B, P, T, D, s, thr) in [(3, 196, 77, 512, 1.0, None), (2, 50, 20, 64, 5.0, None), (2, 196, 77, 512, 14.0, 0.5)]:
        v = torch.randn(B, P, D).to(torch.bfloat16).double(); l = torch.randn(B, T, D).to(torch.bfloat16).double()
        mask = torch.ones(B, T, dtype=torch.bool)
        thr = float(torch.tensor(1.0 / P, dtype=torch.float32)) if thr is None else thr
        f = lo.sparc_forward(v, l, mask, thr, 0.0, 1.0, s)      # local part only: gw = 0
        rv, rl = lo.sparc_backward(f)
        nv = float(f['_cache']['n_valid'])
        z = torch.zeros(B, D, dtype=torch.float64)
        for split in ['none', 'hilo', 'fp16', 'bf16']:
            dv, dl = emul(v, l, mask, thr, s, 0.5 / nv, 0.5 / nv, z, z, split)
            print((B, P, T, D, s), split, 'dv err %.3e  dl err %.3e' % (float((dv - rv).norm() / rv.norm()), float((dl - rl).norm() / rl.norm())))
        # fp16 embeddings (torch.autocast's default): the exact route needs fp16 hi/lo on-chip operands
        v16 = v.to(torch.float16).double(); l16 = l.to(torch.float16).double()
        f = lo.sparc_forward(v16, l16, mask, thr, 0.0, 1.0, s)
        rv, rl = lo.sparc_backward(f)
        for split in ['h16hilo', 'h16hilo_scaled']:
            dv, dl = emul(v16, l16, mask, thr, s, 0.5 / nv, 0.5 / nv, z, z, split)
            print((B, P, T, D, s), 'fp16 inputs', split, 'dv err %.3e  dl err %.3e' % (float((dv - rv).norm() / rv.norm()), float((dl - rl).norm() / rl.norm())))
