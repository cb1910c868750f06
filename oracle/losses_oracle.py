"""CPU oracle for the contrastive-loss hot path.  TEST INFRASTRUCTURE ONLY.

This file is a restatement, in plain torch CPU ops with *hand-derived* backward
formulas (no autograd), of the algorithm in the reference's
``finetune/losses.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product path (``clip_finegrained_alignment_b200``) never does.

Parity pinning: the reference ships no golden vectors or tests (SURVEY.md §4,
§8c), so this oracle is pinned against outputs of the reference itself:
``tests/golden/make_golden.py`` imports ``/root/reference/finetune/losses.py``
in the build container, runs it (autograd for the gradients) on seeded inputs
and stores inputs + outputs in ``tests/golden/*.pt``; ``tests/test_oracle.py``
checks every function below against those fixtures (and against the live
reference when ``/root/reference`` is present).

All functions are dtype-generic (fp32 / fp64); every block cites the reference
``file:line`` it follows (paths relative to ``/root/reference``).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

NORM_EPS = 1e-12      # F.normalize default eps (finetune/losses.py:152-153,173-174,207,212,221-222)
MINMAX_EPS = 1e-8     # finetune/losses.py:231
CLAMP_EPS = 1e-8      # finetune/losses.py:211,242
NVALID_EPS = 1e-8     # finetune/losses.py:196

SPARC_KEYS = ("global_loss", "local_loss", "total_loss", "loss_vl", "loss_lv",
              "loss_vl_local", "loss_lv_local")


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def l2_normalize(x: torch.Tensor, eps: float = NORM_EPS) -> Tuple[torch.Tensor, torch.Tensor]:
    """F.normalize(x, dim=-1): x / max(||x||_2, eps).  Returns (y, clamped norm)."""
    n = x.norm(dim=-1, keepdim=True).clamp_min(eps)
    return x / n, n


def l2_normalize_bwd(y: torch.Tensor, n: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """J_n(x)^T dy = (dy - y (y.dy)) / max(||x||, eps)   (SURVEY.md §8 a-bwd)."""
    return (dy - y * (y * dy).sum(dim=-1, keepdim=True)) / n


def _round_to(x: torch.Tensor, dtype: Optional[torch.dtype]) -> torch.Tensor:
    """Round a contraction operand to `dtype` and come back (matched-rounding oracle-B)."""
    if dtype is None:
        return x
    return x.to(dtype).to(x.dtype)


def _lse(x: torch.Tensor, dim: int) -> torch.Tensor:
    return torch.logsumexp(x, dim=dim)


# --------------------------------------------------------------------------
# InfoNCE on [B, D] x [Bg, D]  (finetune/losses.py:145-163 and :14-36)
# --------------------------------------------------------------------------
def infonce_forward(a: torch.Tensor, b: torch.Tensor, scale: float, col_offset: int = 0,
                    eps: float = NORM_EPS) -> Dict[str, torch.Tensor]:
    """Row-direction InfoNCE of local rows `a` [B,D] against global columns `b` [Bg,D].

    logits = normalize(a) @ normalize(b)^T * scale   (losses.py:152-160)
    row i's target is column `col_offset + i`       (losses.py:162: arange targets)
    Returns the *sum* over local rows of CE (caller divides by the global batch,
    losses.py:163) plus what backward needs.
    """
    ah, an = l2_normalize(a, eps)
    bh, bn = l2_normalize(b, eps)
    logits = ah @ bh.t() * scale
    lse = _lse(logits, dim=1)
    idx = torch.arange(a.shape[0]) + col_offset
    diag = logits[torch.arange(a.shape[0]), idx]
    return {"loss_sum": (lse - diag).sum(), "lse": lse, "logits": logits,
            "ah": ah, "an": an, "bh": bh, "bn": bn}


def clip_loss_forward(img: torch.Tensor, txt: torch.Tensor, temperature: float) -> Dict[str, torch.Tensor]:
    """CustomCLIPLoss.forward (finetune/losses.py:14-36): x/||x|| (no eps), logits / T,
    mean CE on logits and logits.t(), averaged."""
    B = img.shape[0]
    f1 = infonce_forward(img, txt, 1.0 / temperature, eps=0.0)
    f2 = infonce_forward(txt, img, 1.0 / temperature, eps=0.0)
    loss = 0.5 * (f1["loss_sum"] / B + f2["loss_sum"] / B)
    return {"clip_loss": loss, "total_loss": loss, "_f1": f1, "_f2": f2}


def symmetric_infonce_backward(ah, an, bh, bn, lse_a, lse_b, scale: float, c_a: float, c_b: float,
                               denom: float, col_offset: int = 0):
    """Gradient of  c_a * sum_i CE_a(i)/denom + c_b * sum_j CE_b(j)/denom  w.r.t. the
    un-normalised a (rows) and b (rows), single process (a, b both hold all rows).

    dS = (c_a softmax_row(S) + c_b softmax_col(S) - (c_a+c_b) I) / denom ; da = s dS b^ ; db = s dS^T a^
    (SURVEY.md §8 a-bwd, first bullet), then J_n.
    """
    S = ah @ bh.t() * scale
    Pa = torch.exp(S - lse_a[:, None])          # softmax over columns for each a-row
    Pb = torch.exp(S - lse_b[None, :])          # softmax over a-rows for each b-row (column of S)
    dS = c_a * Pa + c_b * Pb
    n = S.shape[0]
    dS[torch.arange(n), torch.arange(n) + col_offset] -= (c_a + c_b)
    dS = dS / denom
    dah = scale * dS @ bh
    dbh = scale * dS.t() @ ah
    return l2_normalize_bwd(ah, an, dah), l2_normalize_bwd(bh, bn, dbh)


def clip_loss_backward(fwd: Dict[str, torch.Tensor], temperature: float, grad_out: float = 1.0):
    f1, f2 = fwd["_f1"], fwd["_f2"]
    B = f1["ah"].shape[0]
    return symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"],
                                      1.0 / temperature, 0.5 * grad_out, 0.5 * grad_out, float(B))


def gathered_infonce_rank(a_loc, b_loc, a_all, b_all, rank: int, scale: float, c_a: float, c_b: float):
    """What ONE rank computes for the all-gathered global InfoNCE (design A).

    Forward: rows = local a vs all b (direction a), rows = local b vs all a (direction b);
    loss = (c_a sum_i CE_a(i) + c_b sum_j CE_b(j)) / Bg summed over all ranks.
    Backward for local rows needs the other direction's LSE for *all* global rows
    (exchanged with a [Bg]-float all-gather); returns a closure computing it.
    """
    B = a_loc.shape[0]
    off = rank * B
    fa = infonce_forward(a_loc, b_all, scale, off)
    fb = infonce_forward(b_loc, a_all, scale, off)

    def backward(lse_a_all, lse_b_all):
        Bg = a_all.shape[0]
        Sa = fa["logits"]                                   # [B, Bg]  local a x all b
        Sb = fb["logits"]                                   # [B, Bg]  local b x all a
        dSa = c_a * torch.exp(Sa - fa["lse"][:, None]) + c_b * torch.exp(Sa - lse_b_all[None, :])
        dSb = c_b * torch.exp(Sb - fb["lse"][:, None]) + c_a * torch.exp(Sb - lse_a_all[None, :])
        ar = torch.arange(B)
        dSa[ar, ar + off] -= (c_a + c_b)
        dSb[ar, ar + off] -= (c_a + c_b)
        dah = scale * (dSa / Bg) @ fa["bh"]
        dbh = scale * (dSb / Bg) @ fb["bh"]
        return l2_normalize_bwd(fa["ah"], fa["an"], dah), l2_normalize_bwd(fb["ah"], fb["an"], dbh)

    return fa, fb, backward


# --------------------------------------------------------------------------
# SPARC  (finetune/losses.py:199-264)
# --------------------------------------------------------------------------
def sparc_forward(v: torch.Tensor, l: torch.Tensor, mask: torch.Tensor, thr: float, gw: float, lw: float,
                  s: float, *, mask_semantics: str = "reference",
                  round_dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
    """SPARCLoss.forward restated.  v [B,P,D], l [B,T,D], mask [B,T] bool.

    mask_semantics:
      "reference": literally what losses.py does (any False in the mask -> NaN local loss,
                   SURVEY.md finding 3).
      "truncate" : masked tokens are skipped (rows/cols contribute nothing); equals the
                   reference evaluated per sample on its valid tokens only, weighted by
                   token count.  This is what the CUDA kernels implement.  Identical to
                   "reference" for all-True masks.
    round_dtype: if set (torch.bfloat16), operands of the contractions that the tensor-core
                 kernels feed as bf16 are rounded first ("matched-rounding" oracle-B, §8c).
    """
    B, P, D = v.shape
    T = l.shape[1]
    mf = mask.to(v.dtype)
    out: Dict[str, torch.Tensor] = {}

    # ---- global pooling (losses.py:207-212)
    vbar_raw = v.mean(dim=1)
    # bool.sum() is int64; clamp(min=1e-8) promotes to fp32, i.e. the count itself when >= 1 (losses.py:211)
    cnt = mask.sum(dim=-1, keepdim=True).clamp(min=CLAMP_EPS).to(v.dtype)
    lbar_raw = (l * mf[..., None]).sum(dim=1) / cnt
    vbar, vbar_n = l2_normalize(vbar_raw)
    lbar, lbar_n = l2_normalize(lbar_raw)
    # ---- global InfoNCE both directions (losses.py:215-217; second normalize at :152-153 is idempotent)
    g1 = infonce_forward(vbar, lbar, s)
    g2 = infonce_forward(lbar, vbar, s)
    loss_vl = g1["loss_sum"] / B
    loss_lv = g2["loss_sum"] / B
    global_loss = 0.5 * (loss_vl + loss_lv)

    # ---- fine-grained similarity (losses.py:221-225)
    vh, vn = l2_normalize(v)
    lh, ln = l2_normalize(l)
    S = torch.einsum("btd,bpd->btp", _round_to(lh, round_dtype), _round_to(vh, round_dtype))
    # ---- masked min-max (losses.py:228-232)
    Sm = S * mf[..., None]
    if mask_semantics == "reference":
        mn = Sm.masked_fill(~mask[..., None], float("inf")).min(dim=-1, keepdim=True)[0]
        mx = Sm.masked_fill(~mask[..., None], -float("inf")).max(dim=-1, keepdim=True)[0]
    else:
        mn = Sm.min(dim=-1, keepdim=True)[0]
        mx = Sm.max(dim=-1, keepdim=True)[0]
    rng = mx - mn + MINMAX_EPS
    N = (Sm - mn) / rng
    # ---- threshold + renorm (losses.py:235-243)
    keep = ~(N < thr)
    Theta = torch.where(keep, N, torch.zeros_like(N))
    sigma = Theta.sum(dim=-1, keepdim=True).clamp_min(CLAMP_EPS)
    W = Theta / sigma
    # ---- language-grouped pooling with RAW v (losses.py:245)
    G = torch.einsum("btp,bpd->btd", _round_to(W, round_dtype), v)
    if mask_semantics == "truncate":
        G = G * mf[..., None]
    # ---- masked token-level InfoNCE both directions (losses.py:165-197, called :248,:250)
    gh, gn = l2_normalize(G)
    L = torch.einsum("bid,bjd->bij", _round_to(gh, round_dtype), _round_to(lh, round_dtype)) * s   # a=G rows, b=l cols
    m2 = mask[:, :, None] & mask[:, None, :]
    Lm = L.masked_fill(~m2, -float("inf"))
    # int64 + 1e-8 promotes to fp32: the eps vanishes for any count >= 1 (losses.py:196)
    n_valid = (mask.sum() + NVALID_EPS).to(v.dtype)
    ar = torch.arange(T)
    lse_r = _lse(Lm, dim=2)                     # vl_local: rows of G vs all tokens
    lse_c = _lse(Lm, dim=1)                     # lv_local: second call's logits are the transpose
    diag = Lm[:, ar, ar]
    if mask_semantics == "truncate":
        ce_r = torch.where(mask, lse_r - diag, torch.zeros_like(diag))
        ce_c = torch.where(mask, lse_c - diag, torch.zeros_like(diag))
    else:
        ce_r = (lse_r - diag) * mf
        ce_c = (lse_c - diag) * mf
    loss_vl_local = ce_r.sum() / n_valid
    loss_lv_local = ce_c.sum() / n_valid
    local_loss = 0.5 * (loss_vl_local + loss_lv_local)
    total = gw * global_loss + lw * local_loss      # losses.py:254

    out.update(global_loss=global_loss, local_loss=local_loss, total_loss=total, loss_vl=loss_vl,
               loss_lv=loss_lv, loss_vl_local=loss_vl_local, loss_lv_local=loss_lv_local)
    out["_cache"] = dict(v=v, l=l, mask=mask, mf=mf, thr=thr, gw=gw, lw=lw, s=s, cnt=cnt,
                         vbar=vbar, vbar_n=vbar_n, lbar=lbar, lbar_n=lbar_n, g1=g1, g2=g2,
                         vh=vh, vn=vn, lh=lh, ln=ln, S=S, mn=mn, mx=mx, rng=rng, N=N, keep=keep,
                         sigma=sigma, W=W, G=G, gh=gh, gn=gn, Lm=Lm, lse_r=lse_r, lse_c=lse_c,
                         n_valid=n_valid, m2=m2, round_dtype=round_dtype, mask_semantics=mask_semantics,
                         vbar_raw=vbar_raw, lbar_raw=lbar_raw)
    return out


def sparc_coefficients(grads: Dict[str, float], gw: float, lw: float) -> Tuple[float, float, float, float]:
    """Fold upstream grads of the 7 outputs (losses.py:256-264) into one coefficient per
    elementary loss: (c_vl, c_lv, c_vl_local, c_lv_local)."""
    g = {k: float(grads.get(k, 0.0)) for k in SPARC_KEYS}
    c_vl = g["loss_vl"] + 0.5 * (g["global_loss"] + gw * g["total_loss"])
    c_lv = g["loss_lv"] + 0.5 * (g["global_loss"] + gw * g["total_loss"])
    c_vll = g["loss_vl_local"] + 0.5 * (g["local_loss"] + lw * g["total_loss"])
    c_lvl = g["loss_lv_local"] + 0.5 * (g["local_loss"] + lw * g["total_loss"])
    return c_vl, c_lv, c_vll, c_lvl


def sparc_backward(fwd: Dict[str, torch.Tensor], grads: Optional[Dict[str, float]] = None):
    """Hand-derived backward of SPARCLoss.forward (SURVEY.md §8 a-bwd); returns (dv, dl).

    `grads` maps output key -> upstream scalar gradient (default: total_loss=1).
    Valid for all-True masks under "reference" semantics and any mask under "truncate".
    """
    c = fwd["_cache"]
    grads = grads or {"total_loss": 1.0}
    c_vl, c_lv, c_vll, c_lvl = sparc_coefficients(grads, c["gw"], c["lw"])
    v, l, mf, mask = c["v"], c["l"], c["mf"], c["mask"]
    B, P, D = v.shape
    T = l.shape[1]
    s = c["s"]
    rd = c["round_dtype"]

    # ---- global InfoNCE (a3): both directions share one logits matrix
    dvbar, dlbar = symmetric_infonce_backward(c["g1"]["ah"], c["g1"]["an"], c["g1"]["bh"], c["g1"]["bn"],
                                              c["g1"]["lse"], c["g2"]["lse"], s, c_vl, c_lv, float(B))
    # through the first normalize (losses.py:207,212) and the pooling (a2)
    dvbar_raw = l2_normalize_bwd(c["vbar"], c["vbar_n"], dvbar)
    dlbar_raw = l2_normalize_bwd(c["lbar"], c["lbar_n"], dlbar)
    dv = (dvbar_raw / P)[:, None, :].expand(B, P, D).clone()
    dl = (dlbar_raw / c["cnt"])[:, None, :] * mf[..., None]

    # ---- masked local CE (a8)
    Lm = c["Lm"]
    Pr = torch.exp(Lm - c["lse_r"][:, :, None])
    Pc = torch.exp(Lm - c["lse_c"][:, None, :])
    Pr = torch.where(c["m2"], Pr, torch.zeros_like(Pr))
    Pc = torch.where(c["m2"], Pc, torch.zeros_like(Pc))
    dL = c_vll * Pr + c_lvl * Pc
    ar = torch.arange(T)
    dL[:, ar, ar] -= (c_vll + c_lvl) * mf
    dL = dL / c["n_valid"]
    dgh = s * torch.einsum("bij,bjd->bid", _round_to(dL, rd), _round_to(c["lh"], rd))
    dlh = s * torch.einsum("bij,bid->bjd", _round_to(dL, rd), _round_to(c["gh"], rd))
    dG = l2_normalize_bwd(c["gh"], c["gn"], dgh)
    if c["mask_semantics"] == "truncate":
        dG = dG * mf[..., None]
    # ---- pooling (a7)
    dv = dv + torch.einsum("btp,btd->bpd", _round_to(c["W"], rd), _round_to(dG, rd))
    dW = torch.einsum("btd,bpd->btp", _round_to(dG, rd), v)
    # ---- renorm + threshold (a6)
    dTheta = (dW - (c["W"] * dW).sum(dim=-1, keepdim=True)) / c["sigma"]
    dN = torch.where(c["keep"], dTheta, torch.zeros_like(dTheta))
    # ---- min-max (a5): scatter to argmin/argmax (first occurrence, like torch.min/max)
    r = c["rng"]
    dSm = dN / r
    dmn = (dN * (c["N"] - 1.0)).sum(dim=-1) / r[..., 0]
    dmx = -(dN * c["N"]).sum(dim=-1) / r[..., 0]
    Sm = c["S"] * mf[..., None]
    imin = Sm.argmin(dim=-1)
    imax = Sm.argmax(dim=-1)
    dSm.scatter_add_(2, imin[..., None], dmn[..., None])
    dSm.scatter_add_(2, imax[..., None], dmx[..., None])
    dS = dSm * mf[..., None]
    # ---- similarity (a4)
    dlh = dlh + torch.einsum("btp,bpd->btd", _round_to(dS, rd), _round_to(c["vh"], rd))
    dvh = torch.einsum("btp,btd->bpd", _round_to(dS, rd), _round_to(c["lh"], rd))
    dl = dl + l2_normalize_bwd(c["lh"], c["ln"], dlh)
    dv = dv + l2_normalize_bwd(c["vh"], c["vn"], dvh)
    return dv, dl


def masked_pairwise_forward(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor, s: float,
                            mask_semantics: str = "truncate") -> Dict[str, torch.Tensor]:
    """SPARCLoss.masked_pairwise_contrastive_loss restated (finetune/losses.py:165-197): a, b [B,T,D], mask [B,T].
    "reference" semantics reproduce the NaN of the reference for padded masks; "truncate" skips masked tokens."""
    T = a.shape[1]
    ah, an = l2_normalize(a)                                             # losses.py:173-174
    bh, bn = l2_normalize(b)
    m2 = mask[:, :, None] & mask[:, None, :]                             # losses.py:177
    L = torch.einsum("bid,bjd->bij", ah, bh) * s                         # losses.py:180
    Lm = L.masked_fill(~m2, -float("inf"))                               # losses.py:186
    lse = _lse(Lm, dim=2)
    ar = torch.arange(T)
    diag = Lm[:, ar, ar]
    mf = mask.to(a.dtype)
    if mask_semantics == "truncate":
        ce = torch.where(mask, lse - diag, torch.zeros_like(diag))
    else:
        ce = (lse - diag) * mf                                           # losses.py:189-196
    n_valid = (mask.sum() + NVALID_EPS).to(a.dtype)
    return {"loss": ce.sum() / n_valid, "_cache": dict(ah=ah, an=an, bh=bh, bn=bn, Lm=Lm, lse=lse, m2=m2, mf=mf,
                                                       n_valid=n_valid, s=s)}


def masked_pairwise_backward(fwd: Dict[str, torch.Tensor], grad_out: float = 1.0):
    c = fwd["_cache"]
    T = c["Lm"].shape[1]
    Pr = torch.where(c["m2"], torch.exp(c["Lm"] - c["lse"][:, :, None]), torch.zeros_like(c["Lm"]))
    dL = Pr.clone()
    ar = torch.arange(T)
    dL[:, ar, ar] -= c["mf"]
    dL = dL * (grad_out / c["n_valid"])
    dah = c["s"] * torch.einsum("bij,bjd->bid", dL, c["bh"])
    dbh = c["s"] * torch.einsum("bij,bid->bjd", dL, c["ah"])
    return l2_normalize_bwd(c["ah"], c["an"], dah), l2_normalize_bwd(c["bh"], c["bn"], dbh)


def sparc_reference_truncated(ref_loss_module, v, l, mask):
    """Per-sample-truncated evaluation of the *reference module* (used to pin the
    "truncate" semantics): local terms are evaluated on each sample's valid tokens only
    and recombined with the batch-wide token count (losses.py:196)."""
    B = v.shape[0]
    tot_vl = 0.0
    tot_lv = 0.0
    n = 0
    for b in range(B):
        idx = mask[b].nonzero()[:, 0]
        if idx.numel() == 0:
            continue
        lb = l[b:b + 1, idx]
        mb = torch.ones(1, idx.numel(), dtype=torch.bool)
        out = ref_loss_module(v[b:b + 1], lb, mb)
        tot_vl = tot_vl + out["loss_vl_local"] * idx.numel()
        tot_lv = tot_lv + out["loss_lv_local"] * idx.numel()
        n += idx.numel()
    n_valid = float(torch.tensor(n) + NVALID_EPS)      # fp32 promotion as in losses.py:196
    return tot_vl / n_valid, tot_lv / n_valid


# --------------------------------------------------------------------------
# Counting losses (finetune/losses.py:39-133 CLIPCountLoss, :267-309 CountLoss) — SURVEY.md §8f rank 3
# --------------------------------------------------------------------------
def logits_ce_forward(la: torch.Tensor, lb: torch.Tensor) -> Dict[str, torch.Tensor]:
    """(CE(la, arange) + CE(lb, arange)) / 2, mean reduction (losses.py:276-279)."""
    B = la.shape[0]
    ar = torch.arange(B)
    lse_a, lse_b = _lse(la, 1), _lse(lb, 1)
    loss = 0.5 * ((lse_a - la[ar, ar]).mean() + (lse_b - lb[ar, ar]).mean())
    return {"loss": loss, "lse_a": lse_a, "lse_b": lse_b}


def logits_ce_backward(la, lb, fwd, grad_out: float = 1.0):
    B = la.shape[0]
    eye = torch.eye(B, dtype=la.dtype)
    c = grad_out * 0.5 / B
    return c * (torch.exp(la - fwd["lse_a"][:, None]) - eye), c * (torch.exp(lb - fwd["lse_b"][:, None]) - eye)


def count_contrastive_forward(ei, ek, ek_cf, temperature: float, include_pos: bool = False):
    """losses.py:281-301: rows normalised without eps; loss_b = -log(exp(pos) / sum_c exp(cf_c)), mean over b.
    include_pos adds exp(pos) to the denominator (the grouping of CLIPCountLoss.count_loss, losses.py:78-86)."""
    ih, inn = l2_normalize(ei, 0.0)
    kh, kn = l2_normalize(ek, 0.0)
    ch, cn = l2_normalize(ek_cf, 0.0)
    pos = (ih * kh).sum(1) / temperature                        # :289
    cfs = (ih[:, None, :] * ch).sum(2) / temperature            # :297
    allc = torch.cat([pos[:, None], cfs], dim=1) if include_pos else cfs
    lse = _lse(allc, 1)
    loss = (lse - pos).mean()                                   # :301-303
    return {"loss": loss, "_c": dict(ih=ih, inn=inn, kh=kh, kn=kn, ch=ch, cn=cn, pos=pos, cfs=cfs, lse=lse,
                                     T=temperature, include_pos=include_pos)}


def count_contrastive_backward(fwd, grad_out: float = 1.0):
    c = fwd["_c"]
    B = c["pos"].shape[0]
    gb = grad_out / B
    dcf = gb * torch.exp(c["cfs"] - c["lse"][:, None])          # d loss / d cf score
    dpos = torch.full_like(c["pos"], -gb)
    if c["include_pos"]:
        dpos = dpos + gb * torch.exp(c["pos"] - c["lse"])
    dih = (dpos[:, None] * c["kh"] + (dcf[:, :, None] * c["ch"]).sum(1)) / c["T"]
    dkh = dpos[:, None] * c["ih"] / c["T"]
    dch = dcf[:, :, None] * c["ih"][:, None, :] / c["T"]
    return (l2_normalize_bwd(c["ih"], c["inn"], dih), l2_normalize_bwd(c["kh"], c["kn"], dkh),
            l2_normalize_bwd(c["ch"], c["cn"], dch))


def count_loss_forward(img_logits, text_logits, ei, ek, ek_cf, temperature: float, alpha: float):
    """CountLoss.forward (losses.py:273-309)."""
    f1 = logits_ce_forward(img_logits, text_logits)
    f2 = count_contrastive_forward(ei, ek, ek_cf, temperature)
    return {"clip_loss": f1["loss"], "count_loss": f2["loss"], "total_loss": f1["loss"] + alpha * f2["loss"],
            "_f1": f1, "_f2": f2}


def clip_count_forward(img, txt, temperature: float):
    """CLIPCountLoss.forward (losses.py:91-133) for one count per caption: the CLIP loss on the template-expanded
    batch (image rows repeated, :104); the count term is exactly 0 (group of one, :69-86)."""
    nt = txt.shape[0] // img.shape[0]
    f = clip_loss_forward(img.repeat_interleave(nt, dim=0), txt, temperature)
    return {"clip_loss": f["clip_loss"], "count_loss": torch.zeros((), dtype=torch.float64), "total_loss": f["clip_loss"],
            "_f": f, "_nt": nt}


def clip_count_backward(fwd, temperature: float, grad_out: float = 1.0):
    da_exp, db = clip_loss_backward(fwd["_f"], temperature, grad_out)
    nt = fwd["_nt"]
    return da_exp.view(-1, nt, da_exp.shape[1]).sum(1), db      # repeat_interleave backward: sum over the copies
