"""SPARCLoss and CustomCLIPLoss with the reference's API (finetune/losses.py), computed by the
sm_100a kernels behind include/cfa_b200.h.

    SPARCLoss(config)(v_patch_embed[B,P,D], l_token_embed[B,T,D], language_mask[B,T] bool) -> dict of 7
    CustomCLIPLoss(temperature)(image_features[B,D], text_features[B,D], custom_features=None) -> dict of 2

Every returned entry is a 0-dim CUDA tensor attached to autograd; the backward is hand-written
(kernels), not traced.  Inputs may be fp32, bf16 or fp16 (what autocast hands the reference,
finetuner.py:120-134); gradients come back in the input dtype.

Beyond the reference: `gather=True` (or an explicit `process_group`) turns the *global* InfoNCE into
the all-gathered variant — image/text embeddings are all-gathered with NCCL and every rank scores its
local rows against the global columns.  The fine-grained loss stays rank-local, as in
dist_finetuner.py.

Padded masks: the reference's local loss is NaN as soon as one mask entry is False (SURVEY finding 3).
The kernels implement the evidently intended semantics instead ("truncate": masked tokens are skipped,
exactly the reference evaluated per sample on its valid tokens); for all-True masks — the only case
the reference can train with — results are identical.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

_L = _lib.lib
SPARC_KEYS = ("global_loss", "local_loss", "total_loss", "loss_vl", "loss_lv", "loss_vl_local", "loss_lv_local")
_NORM_EPS = 1e-12


# ----------------------------------------------------------------------------------------------
# thin wrappers over the C ABI (allocation happens here; the library never allocates)
# ----------------------------------------------------------------------------------------------
def _rows_normalize(x: torch.Tensor, eps: float):
    xh = torch.empty_like(x)
    n = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _lib.call("cfa_rows_normalize", x.data_ptr(), x.shape[0], x.shape[1], eps, xh.data_ptr(), n.data_ptr(),
                                     _lib.stream_ptr())
    return xh, n


def _infonce_fwd(a_hat: torch.Tensor, b_hat_all: torch.Tensor, col_offset: int, scale: float):
    B, D = a_hat.shape
    Bg = b_hat_all.shape[0]
    ws_bytes = _L.cfa_infonce_fwd_workspace_bytes(B, Bg, D)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a_hat.device)
    lse = torch.empty(B, dtype=torch.float32, device=a_hat.device)
    ce = torch.empty(B, dtype=torch.float32, device=a_hat.device)
    _lib.call("cfa_infonce_fwd", a_hat.data_ptr(), B, b_hat_all.data_ptr(), Bg, D, col_offset, scale, lse.data_ptr(),
                                  ce.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr())
    return lse, ce


def _infonce_bwd(a_hat, a_norm, b_hat_all, col_offset, scale, lse_a, lse_b_all, coef):
    """Gradient w.r.t. the UN-normalised local rows `a` (J_n applied)."""
    import ctypes
    B, D = a_hat.shape
    Bg = b_hat_all.shape[0]
    npart = ctypes.c_int(0)
    ws_bytes = _L.cfa_infonce_bwd_workspace_bytes(B, Bg, D, ctypes.byref(npart))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a_hat.device)
    _lib.call("cfa_infonce_bwd", a_hat.data_ptr(), B, b_hat_all.data_ptr(), Bg, D, col_offset, scale, lse_a.data_ptr(),
                                  lse_b_all.data_ptr(), coef.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr())
    da = torch.empty_like(a_hat)
    _lib.call("cfa_rows_normalize_bwd", a_hat.data_ptr(), a_norm.data_ptr(), ws.data_ptr(), npart.value, B * D, B, D,
                                         da.data_ptr(), _lib.stream_ptr())
    return da


def _sum2(x0, x1):
    out = torch.empty(2, dtype=torch.float32, device=x0.device)
    _lib.call("cfa_sum2", x0.data_ptr(), x1.data_ptr(), x0.numel(), out.data_ptr(), _lib.stream_ptr())
    return out


def _dist_ctx(group, gather: bool):
    """(world, rank, group) of the gathered global loss; world == 1 means rank-local."""
    if not gather:
        return 1, 0, None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0, None
    return dist.get_world_size(group), dist.get_rank(group), group


def _all_gather_rows(x: torch.Tensor, world: int, group) -> torch.Tensor:
    """NCCL all-gather of equally sized row blocks (the one exchange step of the path)."""
    if world == 1:
        return x
    import torch.distributed as dist
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


# ----------------------------------------------------------------------------------------------
# global InfoNCE on [B, D] embeddings: forward state + backward, shared by both losses
# ----------------------------------------------------------------------------------------------
class _GlobalState:
    __slots__ = ("ah", "an", "bh", "bn", "ah_all", "bh_all", "lse_a", "lse_b", "off", "world", "group", "scale", "Bg")


def _global_forward(a: torch.Tensor, b: torch.Tensor, scale: float, eps: float, world: int, rank: int, group):
    """a, b: local fp32 [B, D].  Returns (state, sums[2]) with sums = (sum_i CE_a(i), sum_j CE_b(j)) over
    the GLOBAL batch (all-reduced when world > 1)."""
    st = _GlobalState()
    st.world, st.group, st.scale = world, group, scale
    B = a.shape[0]
    st.off = rank * B
    st.Bg = world * B
    st.ah, st.an = _rows_normalize(a, eps)
    st.bh, st.bn = _rows_normalize(b, eps)
    st.ah_all = _all_gather_rows(st.ah, world, group)
    st.bh_all = _all_gather_rows(st.bh, world, group)
    st.lse_a, ce_a = _infonce_fwd(st.ah, st.bh_all, st.off, scale)     # image rows vs all text columns
    st.lse_b, ce_b = _infonce_fwd(st.bh, st.ah_all, st.off, scale)     # text rows vs all image columns
    sums = _sum2(ce_a, ce_b)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(sums, group=group)
    return st, sums


def _global_backward(st: _GlobalState, coef_a: torch.Tensor, coef_b: torch.Tensor):
    """coef_a = device [2] (c_a, c_b)/Bg, coef_b = (c_b, c_a)/Bg.  Returns (da, db) w.r.t. the local
    un-normalised rows; cross-rank terms come from the other direction's gathered LSE vector."""
    lse_a_all = _all_gather_rows(st.lse_a, st.world, st.group)
    lse_b_all = _all_gather_rows(st.lse_b, st.world, st.group)
    da = _infonce_bwd(st.ah, st.an, st.bh_all, st.off, st.scale, st.lse_a, lse_b_all, coef_a)
    db = _infonce_bwd(st.bh, st.bn, st.ah_all, st.off, st.scale, st.lse_b, lse_a_all, coef_b)
    return da, db


# ----------------------------------------------------------------------------------------------
# SPARC
# ----------------------------------------------------------------------------------------------
class _SparcFunction(torch.autograd.Function):
    """Returns one [7] vector (SPARC_KEYS order); the module hands out its elements, so autograd delivers
    a single [7] gradient vector to backward — no host synchronisation anywhere."""

    @staticmethod
    def forward(ctx, v, l, mask, thr, gw, lw, scale, gather, group, path):
        dev = _lib.require_cuda(v, l, mask)
        if v.dtype != l.dtype or v.dtype not in _lib.DTYPE_CODE:
            raise _lib.CfaError(f"SPARCLoss: embeddings must share a dtype in fp32/bf16/fp16, got {v.dtype}, {l.dtype}")
        if v.dim() != 3 or l.dim() != 3 or mask.dim() != 2 or v.shape[0] != l.shape[0] or v.shape[2] != l.shape[2] \
                or mask.shape != l.shape[:2]:
            raise _lib.CfaError(f"SPARCLoss: bad shapes v{tuple(v.shape)} l{tuple(l.shape)} mask{tuple(mask.shape)}")
        v = v.contiguous()
        l = l.contiguous()
        mask_u8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else (mask != 0).contiguous().view(torch.uint8)
        B, P, D = v.shape
        T = l.shape[1]
        code = _lib.DTYPE_CODE[v.dtype]
        f32 = dict(dtype=torch.float32, device=dev)
        pooled_v = torch.empty(B, D, **f32)
        pooled_l = torch.empty(B, D, **f32)
        lse_r = torch.empty(B, T, **f32)
        lse_c = torch.empty(B, T, **f32)
        part = torch.empty(B, 2, **f32)
        inv_norm = torch.empty(B * (P + T), **f32)
        with torch.cuda.device(dev):
            _lib.call("cfa_sparc_fwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr, scale,
                      inv_norm.data_ptr(), pooled_v.data_ptr(), pooled_l.data_ptr(), lse_r.data_ptr(), lse_c.data_ptr(),
                      part.data_ptr(), path, _lib.stream_ptr())
            world, rank, group = _dist_ctx(group, gather)
            gst, sums = _global_forward(pooled_v, pooled_l, scale, _NORM_EPS, world, rank, group)
            out8 = torch.empty(8, **f32)
            _lib.call("cfa_sparc_finalize", sums.data_ptr(), gst.Bg, part.data_ptr(), mask_u8.data_ptr(), B, T, gw, lw,
                                             out8.data_ptr(), _lib.stream_ptr())
        ctx.save_for_backward(v, l, mask_u8, lse_r, lse_c, out8, inv_norm)
        ctx.gst = gst
        ctx.hp = (thr, gw, lw, scale, code, path)
        return out8[:7].clone()

    @staticmethod
    def backward(ctx, grad7):
        v, l, mask_u8, lse_r, lse_c, out8, inv_norm = ctx.saved_tensors
        thr, gw, lw, scale, code, path = ctx.hp
        gst = ctx.gst
        B, P, D = v.shape
        T = l.shape[1]
        dev = v.device
        grad7 = grad7.to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            coef = torch.empty(8, dtype=torch.float32, device=dev)
            _lib.call("cfa_sparc_coef", grad7.data_ptr(), gw, lw, gst.Bg, out8.data_ptr(), coef.data_ptr(),
                                         _lib.stream_ptr())
            dpv, dpl = _global_backward(gst, coef[0:2], coef[4:6])
            dv = torch.empty_like(v)
            dl = torch.empty_like(l)
            _lib.call("cfa_sparc_bwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr, scale,
                      inv_norm.data_ptr(), lse_r.data_ptr(), lse_c.data_ptr(), coef[2:4].data_ptr(), dpv.data_ptr(),
                      dpl.data_ptr(), dv.data_ptr(), dl.data_ptr(), path, _lib.stream_ptr())
        return dv, dl, None, None, None, None, None, None, None, None


class _PairwiseFunction(torch.autograd.Function):
    """One-direction InfoNCE of a[B,D] vs b[B,D] (SPARCLoss.pairwise_contrastive_loss, losses.py:145-163)."""

    @staticmethod
    def forward(ctx, a, b, scale, eps):
        dev = _lib.require_cuda(a, b)
        a32 = a.detach().to(torch.float32).contiguous()
        b32 = b.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            ah, an = _rows_normalize(a32, eps)
            bh, bn = _rows_normalize(b32, eps)
            lse_a, ce_a = _infonce_fwd(ah, bh, 0, scale)
            sums = _sum2(ce_a, ce_a)
        ctx.save_for_backward(ah, an, bh, bn, lse_a)
        ctx.meta = (scale, a.dtype, b.dtype)
        return sums[0] / a.shape[0]

    @staticmethod
    def backward(ctx, g):
        ah, an, bh, bn, lse_a = ctx.saved_tensors
        scale, adt, bdt = ctx.meta
        B = ah.shape[0]
        with torch.cuda.device(ah.device):
            z = torch.zeros(1, dtype=torch.float32, device=ah.device)
            c = (g.to(torch.float32).reshape(1) / B)
            coef_a = torch.cat([c, z])                 # rows of a: own-direction term only
            coef_b = torch.cat([z, c])                 # rows of b see it through the columns
            lse_dummy = torch.full((B,), 1e30, dtype=torch.float32, device=ah.device)   # exp(S - 1e30) == 0
            da = _infonce_bwd(ah, an, bh, 0, scale, lse_a, lse_dummy, coef_a)
            db = _infonce_bwd(bh, bn, ah, 0, scale, lse_dummy, lse_a, coef_b)
        return da.to(adt), db.to(bdt), None, None


class SPARCLoss(nn.Module):
    """SPARC loss (https://arxiv.org/abs/2401.09865), reference API: finetune/losses.py:136-264."""

    def __init__(self, config, gather: bool = False, process_group=None, kernel_path: str = "auto"):
        super().__init__()
        # "auto": tcgen05 tensor-core kernels for bf16 inputs of supported shapes, fp32-exact CUDA-core kernels
        # otherwise; "simt" / "tc" force one of them (tests, benchmarks)
        self.kernel_path = {"auto": 0, "simt": 1, "tc": 2}[kernel_path]
        self.similarity_threshold = config.similarity_threshold      # losses.py:140-143
        self.global_loss_weight = config.global_loss_weight
        self.local_loss_weight = config.local_loss_weight
        self.inverse_temperature = config.inverse_temperature
        self.gather = gather
        self.process_group = process_group

    def pairwise_contrastive_loss(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        """a, b: [batch, dim] -> CE(sum)/batch of normalize(a) @ normalize(b).T * inverse_temperature."""
        return _PairwiseFunction.apply(a, b, float(self.inverse_temperature), _NORM_EPS)

    def masked_pairwise_contrastive_loss(self, a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "masked_pairwise_contrastive_loss is fused into SPARCLoss.forward (cfa_sparc_fwd/bwd); a standalone "
            "[B,T,D]x[B,T,D] entry point is not exported yet")

    def forward(self, v_patch_embed: torch.Tensor, l_token_embed: torch.Tensor,
                language_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        if language_mask.dtype not in (torch.bool, torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            # the reference fails in `~language_mask` for float masks (SURVEY §8b errors)
            raise TypeError(f"language_mask must be bool or integer, got {language_mask.dtype}")
        out = _SparcFunction.apply(v_patch_embed, l_token_embed, language_mask, float(self.similarity_threshold),
                                   float(self.global_loss_weight), float(self.local_loss_weight),
                                   float(self.inverse_temperature), self.gather, self.process_group, self.kernel_path)
        return {k: out[i] for i, k in enumerate(SPARC_KEYS)}


# ----------------------------------------------------------------------------------------------
# CLIP InfoNCE
# ----------------------------------------------------------------------------------------------
class _ClipFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, temperature, gather, group):
        dev = _lib.require_cuda(img, txt)
        if img.dim() != 2 or img.shape != txt.shape:
            raise _lib.CfaError(f"CustomCLIPLoss: expected two [B,D] tensors, got {tuple(img.shape)}, {tuple(txt.shape)}")
        a = img.detach().to(torch.float32).contiguous()
        b = txt.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            world, rank, group = _dist_ctx(group, gather)
            # x / x.norm(): no eps in this loss (losses.py:17-18)
            gst, sums = _global_forward(a, b, 1.0 / temperature, 0.0, world, rank, group)
            loss = (sums[0] + sums[1]) * (0.5 / gst.Bg)            # mean CE both ways, averaged (losses.py:27-29)
        ctx.gst = gst
        ctx.dt = (img.dtype, txt.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        gst = ctx.gst
        with torch.cuda.device(gst.ah.device):
            c = (g.to(torch.float32).reshape(1) * (0.5 / gst.Bg)).expand(2).contiguous()
            da, db = _global_backward(gst, c, c)
        return da.to(ctx.dt[0]), db.to(ctx.dt[1]), None, None, None


class CustomCLIPLoss(nn.Module):
    """Symmetric CLIP InfoNCE, reference API: finetune/losses.py:7-36 (logits are DIVIDED by temperature)."""

    def __init__(self, temperature: float = 0.07, gather: bool = False, process_group=None):
        super().__init__()
        self.temperature = temperature
        self.gather = gather
        self.process_group = process_group

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor,
                custom_features: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        clip_loss = _ClipFunction.apply(image_features, text_features, float(self.temperature), self.gather,
                                        self.process_group)
        return {"clip_loss": clip_loss, "total_loss": clip_loss}
