"""SPARCLoss and CustomCLIPLoss with the reference's API (finetune/losses.py), computed by the
sm_100a kernels behind include/cfa_b200.h.

    SPARCLoss(config)(v_patch_embed[B,P,D], l_token_embed[B,T,D], language_mask[B,T] bool) -> dict of 7
    CustomCLIPLoss(temperature)(image_features[B,D], text_features[B,D], custom_features=None) -> dict of 2

Every returned entry is a 0-dim CUDA tensor attached to autograd; the backward is hand-written
(kernels), not traced.  Inputs may be fp32, bf16 or fp16 (what autocast hands the reference,
finetuner.py:120-134); gradients come back in the input dtype.

Beyond the reference: `gather=True` (or an explicit `process_group`) turns the *global* InfoNCE into
the all-gathered variant — every rank scores its local rows against the rows of all ranks (SURVEY.md
§8e).  On one NVLink/NVSwitch box the kernels read the peers' pooled embeddings in place through
CUDA-IPC-mapped exchange blocks (two in-stream device barriers per step, no collective call);
`gather="nccl"` (or a shape the tensor-core global kernels do not take) all-gathers with NCCL instead.
The fine-grained loss stays rank-local, as in dist_finetuner.py.

Padded masks: the reference's local loss is NaN as soon as one mask entry is False (SURVEY finding 3).
The kernels implement the evidently intended semantics instead ("truncate": masked tokens are skipped,
exactly the reference evaluated per sample on its valid tokens); for all-True masks — the only case
the reference can train with — results are identical.
"""
from __future__ import annotations

from typing import Dict, Optional

import contextlib

import torch
import torch.nn as nn

from . import _lib

_L = _lib.lib
SPARC_KEYS = ("global_loss", "local_loss", "total_loss", "loss_vl", "loss_lv", "loss_vl_local", "loss_lv_local")
_NORM_EPS = 1e-12


# ----------------------------------------------------------------------------------------------
# global InfoNCE on [B, D] embeddings (both directions per launch), shared by both losses
# ----------------------------------------------------------------------------------------------
def _dist_ctx(group, gather: bool):
    """(world, rank, group) of the gathered global loss; world == 1 means rank-local."""
    if not gather:
        return 1, 0, None
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0, None
    return dist.get_world_size(group), dist.get_rank(group), group


def _gather_embeddings(ab: torch.Tensor, world: int, group, raw: bool):
    """Collective 1 of 2: all-gather of the [2, B, D] (image | text) local rows.  raw=True (tensor-core path): returns the
    NCCL output [world, 2, B, D] untouched — the kernels index it as it is, no re-layout launch; raw=False: two
    contiguous [Bg, D] matrices in rank-major global row order.  NCCL over NVLink on the GPU box; gloo on CPU in
    tests/test_dist_gloo.py."""
    if world == 1:
        return ab[0], ab[1]
    import torch.distributed as dist
    _, B, D = ab.shape
    gathered = torch.empty(world * 2, B, D, dtype=ab.dtype, device=ab.device)       # concatenated along dim 0
    dist.all_gather_into_tensor(gathered, ab.contiguous(), group=group)
    if raw:
        return gathered[0], gathered[1]            # rank 0's image / text blocks; rank r's sit 2*B*D*r elements further
    ab_all = gathered.view(world, 2, B, D).permute(1, 0, 2, 3).reshape(2, world * B, D).contiguous()
    return ab_all[0], ab_all[1]


def _gather_lse_and_sums(pack: torch.Tensor, B: int, world: int, group, raw: bool):
    """Collective 2 of 2: all-gather of [lse_a | lse_b | sum CE_a, sum CE_b] (2B+2 floats per rank).  raw=True: returns
    the gathered [world, 2B+2] buffer twice (the kernels read lse and sums straight out of it); raw=False:
    (lse_all [2, Bg] rank-major, sums2 [2] over the global batch).  The backward needs nothing else."""
    if world == 1:
        return pack[:2 * B].view(2, B), pack[2 * B:]
    import torch.distributed as dist
    g = torch.empty(world * (2 * B + 2), dtype=pack.dtype, device=pack.device)
    dist.all_gather_into_tensor(g, pack.contiguous(), group=group)
    if raw:
        return g, g
    g = g.view(world, 2 * B + 2)
    sums2 = g[:, 2 * B:].sum(dim=0)
    lse_all = g[:, :2 * B].view(world, 2, B).permute(1, 0, 2).reshape(2, world * B).contiguous()
    return lse_all, sums2


class _GlobalState:
    __slots__ = ("a", "b", "a_all", "b_all", "lse2", "lse_all", "norms2", "off", "world", "group", "scale", "eps", "Bg", "ws",
                 "path", "ranks")


def _global_forward(ab, scale, eps, world, rank, group, fused=None, path=0, raw_sums=False):
    """ab: local RAW fp32 [2, B, D] (image rows, text rows).  Returns (state, sums2) with sums2 = (sum_i CE_a,
    sum_j CE_b) over the GLOBAL batch.  Two collectives per step when world > 1: one all-gather of the embeddings
    and one all-gather of [lse | CE sums] (which also serves the backward, so the backward has no collective).
    `fused` = (local_partial, mask_u8, T, gw, lw, out8) fuses the SPARC scalar epilogue into the same call when the
    loss is rank-local."""
    st = _GlobalState()
    _, B, D = ab.shape
    dev = ab.device
    st.world, st.group, st.scale, st.eps, st.path = world, group, scale, eps, path
    st.off, st.Bg = rank * B, world * B
    st.a, st.b = ab[0], ab[1]
    # raw all-gather buffers are consumed as they are when the tensor-core kernels run (cfa_global_infonce_path == 2)
    raw = world > 1 and ab.is_cuda and _L.cfa_global_infonce_path(B, world * B, D, path) == 2
    st.ranks = world if raw else 0
    st.a_all, st.b_all = _gather_embeddings(ab, world, group, raw)
    # one allocation: [lse_a | lse_b | sum CE_a, sum CE_b] (= one collective message) | norms2 | kernel workspace
    ws_bytes = _L.cfa_global_infonce_workspace_bytes(B, st.Bg, D)
    nf = 4 * B + 4
    buf = torch.empty(nf + (ws_bytes + 3) // 4, dtype=torch.float32, device=dev)
    pack = buf[:2 * B + 2]
    st.lse2 = pack[:2 * B].view(2, B)
    sums2 = pack[2 * B:]
    st.norms2 = buf[2 * B + 4:4 * B + 4].view(2, B)
    st.ws = buf[nf:]
    if fused is not None and world == 1:
        part, mask_u8, T, gw, lw, out8 = fused
        _lib.call("cfa_global_infonce_fwd", st.a.data_ptr(), st.b.data_ptr(), st.a_all.data_ptr(), st.b_all.data_ptr(), B,
                  st.Bg, D, st.off, scale, eps, st.lse2.data_ptr(), st.norms2.data_ptr(), sums2.data_ptr(), part.data_ptr(),
                  mask_u8.data_ptr(), T, gw, lw, out8.data_ptr(), st.ws.data_ptr(), st.ws.numel() * 4, path, 0,
                  _lib.stream_ptr())
        st.lse_all = st.lse2
    else:
        _lib.call("cfa_global_infonce_fwd", st.a.data_ptr(), st.b.data_ptr(), st.a_all.data_ptr(), st.b_all.data_ptr(), B,
                  st.Bg, D, st.off, scale, eps, st.lse2.data_ptr(), st.norms2.data_ptr(), sums2.data_ptr(), 0, 0, 0, 0.0,
                  0.0, 0, st.ws.data_ptr(), st.ws.numel() * 4, path, st.ranks, _lib.stream_ptr())
        st.lse_all, sums2 = _gather_lse_and_sums(pack, B, world, group, raw)
        if raw and not raw_sums:                       # caller wants the two global CE sums as a [2] tensor
            sums2 = sums2.view(world, 2 * B + 2)[:, 2 * B:].sum(dim=0)
    return st, sums2


def _global_backward(st: _GlobalState, coef2: torch.Tensor):
    """coef2 = device [2] (c_a, c_b)/Bg.  Returns (da, db) w.r.t. the local raw rows; cross-rank terms come from
    the other direction's LSE vector gathered in the forward — no collective here."""
    B, D = st.a.shape
    dab = torch.empty(2, B, D, dtype=torch.float32, device=st.a.device)
    _lib.call("cfa_global_infonce_bwd", st.a.data_ptr(), st.b.data_ptr(), st.a_all.data_ptr(), st.b_all.data_ptr(), B,
              st.Bg, D, st.off, st.scale, st.eps, st.lse2.data_ptr(), st.lse_all.data_ptr(), st.norms2.data_ptr(),
              coef2.data_ptr(), dab.data_ptr(), dab.data_ptr() + 4 * B * D, st.ws.data_ptr(), st.ws.numel() * 4,
              st.path, st.ranks, _lib.stream_ptr())
    return dab[0], dab[1]


# ----------------------------------------------------------------------------------------------
# SPARC
# ----------------------------------------------------------------------------------------------
_WS_BYTES = {}      # (B, P, T, D, dtype, path) -> cfa_sparc_loss_workspace_bytes


class _SparcFunction(torch.autograd.Function):
    """Returns the 7 losses (SPARC_KEYS order) as separate 0-dim outputs of ONE autograd node; backward receives
    the 7 upstream gradients (None for unused outputs) and hands their device pointers to the coefficient kernel —
    no host synchronisation, zero-fill or concatenation anywhere."""

    # No torch.amp.custom_fwd / custom_bwd here: the function runs no autocast-sensitive torch op (everything numerical
    # happens inside the library on the dtype the caller hands over, finetuner.py:120), and the two decorators cost ~50 us
    # of host time per step (measured: host issue 0.19 -> 0.24 ms).  tests/test_gpu_trainer.py runs the loss under
    # torch.autocast(fp16 / bf16) + GradScaler.
    @staticmethod
    def forward(ctx, v, l, mask, thr, gw, lw, scale, gather, group, path, fused=True, ddp_mean=True):
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
        world, rank, group = _dist_ctx(group, gather)
        ctx.set_materialize_grads(False)               # unused outputs arrive as None: no zero-fill launches
        ctx.peer = None
        ctx.gscale = float(world) if (ddp_mean and world > 1) else 1.0     # see SPARCLoss.__init__ (gather_grad_reduce)
        if world > 1 and fused and gather != "nccl":
            # gathered loss over peer memory: one library call per direction, no collective call (csrc/peer_exchange.cu)
            gpath = 1 if (v.dtype == torch.float32 or path == 1) else 0
            ex = None
            if _L.cfa_global_infonce_path(B, world * B, D, gpath) == 2:
                from . import peer as _peer
                with torch.cuda.device(dev):
                    ex = _peer.get_exchange(B, D, group)
            if ex is not None:
                key = (B, P, T, D, code, path, world)
                nbytes = _WS_BYTES.get(key)
                if nbytes is None:
                    nbytes = _WS_BYTES[key] = _L.cfa_sparc_loss_gathered_workspace_bytes(B, P, T, D, code, path, world)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                same_dev = torch.cuda.current_device() == dev.index
                with (contextlib.nullcontext() if same_dev else torch.cuda.device(dev)):
                    _lib.call("cfa_sparc_loss_gathered_fwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code,
                              thr, scale, gw, lw, ws.data_ptr(), nbytes, path, world, rank, ex.blocks, ex.next_step(),
                              _lib.stream_ptr())
                ctx.save_for_backward(v, l, mask_u8, ws)
                ctx.gst = None
                ctx.peer = (world, rank)
                ctx.hp = (thr, gw, lw, scale, code, path, None, None)
                return ws[:28].view(torch.float32).clone().unbind(0)
        if world == 1 and fused:
            # rank-local loss: ONE library call and ONE allocation per direction (cfa_sparc_loss_fwd / _bwd); the
            # workspace layout is private to the library, its first 8 floats are the outputs
            key = (B, P, T, D, code, path)
            nbytes = _WS_BYTES.get(key)
            if nbytes is None:
                nbytes = _WS_BYTES[key] = _L.cfa_sparc_loss_workspace_bytes(B, P, T, D, code, path)
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            same_dev = torch.cuda.current_device() == dev.index
            with (contextlib.nullcontext() if same_dev else torch.cuda.device(dev)):
                _lib.call("cfa_sparc_loss_fwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr, scale,
                          gw, lw, ws.data_ptr(), nbytes, path, _lib.stream_ptr())
            ctx.save_for_backward(v, l, mask_u8, ws)
            ctx.gst = None
            ctx.hp = (thr, gw, lw, scale, code, path, None, None)
            return ws[:28].view(torch.float32).clone().unbind(0)
        # ONE fp32 allocation, carved by pointer arithmetic (host time matters at ~0.5 ms per step):
        # pooled [2,B,D] | out8 | lse_row [B,T] | lse_col [B,T] | local_partial [B,2] | row_inv_norm [B(P+T)] |
        # tt_logits [B,T,T] | g_inv_norm [B,T]
        # | g_split [B,2,T,D] bf16 | q_save [B,T,NP]   (every piece starts 128-byte aligned: TMA / float4 access)
        NP = (P + 15) & ~15
        saved = path != 1 and _L.cfa_sparc_path(P, T, D, code, path) == 2      # tensor-core path saves G (hi|lo) and Q
        sizes = (2 * B * D, 8, B * T, B * T, 2 * B, B * (P + T), B * T * T, B * T,
                 B * T * D if saved else 0, B * T * NP if saved else 0)
        off = [0]
        for n in sizes:
            off.append(off[-1] + ((n + 31) & ~31))
        blk = torch.empty(off[-1], dtype=torch.float32, device=dev)
        base = blk.data_ptr()
        ptr = [base + 4 * o for o in off]
        gq = (ptr[8], ptr[9]) if saved else (0, 0)
        # CUDA-core path with P too large for shared-memory residency (ViT-L/14@336): L2-resident global scratch
        tc = path != 1 and _L.cfa_sparc_path(P, T, D, code, path) == 2
        sbytes = 0 if tc else _L.cfa_sparc_scratch_bytes(B, P, T, 1)
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=dev) if sbytes else None
        sptr = scratch.data_ptr() if sbytes else 0
        pooled = blk[:2 * B * D].view(2, B, D)
        out8 = blk[off[1]:off[2]]
        part_t = blk[off[4]:off[5]]
        same_dev = torch.cuda.current_device() == dev.index
        with (contextlib.nullcontext() if same_dev else torch.cuda.device(dev)):
            _lib.call("cfa_sparc_fwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr, scale,
                      ptr[5], ptr[0], ptr[0] + 4 * B * D, ptr[2], ptr[3], ptr[4], ptr[6], ptr[7], gq[0], gq[1], sptr, sbytes,
                      path, _lib.stream_ptr())
            gpath = 1 if (v.dtype == torch.float32 or path == 1) else 0      # fp32 inputs keep the fp32-exact global kernels
            gst, sums = _global_forward(pooled, scale, _NORM_EPS, world, rank, group,
                                        fused=(part_t, mask_u8, T, gw, lw, out8), path=gpath, raw_sums=True)
            if world > 1:       # scalar epilogue after the cross-rank gather of the CE sums
                _lib.call("cfa_sparc_finalize", sums.data_ptr(), gst.Bg, ptr[4], mask_u8.data_ptr(), B, T, gw, lw, ptr[1],
                          gst.ranks, _lib.stream_ptr())
        ctx.save_for_backward(v, l, mask_u8, blk)
        ctx.gst = gst
        ctx.hp = (thr, gw, lw, scale, code, path, ptr, gq)
        ctx.scratch = scratch                          # reused by the backward (same size class)
        return out8[:7].clone().unbind(0)

    @staticmethod
    def backward(ctx, *grads):
        v, l, mask_u8, blk = ctx.saved_tensors
        thr, gw, lw, scale, code, path, ptr, gq = ctx.hp
        gst = ctx.gst
        B, P, D = v.shape
        T = l.shape[1]
        dev = v.device
        gptr = []
        keep = []
        for gk in grads:
            if gk is None:
                gptr.append(0)
                continue
            if gk.dtype != torch.float32 or gk.device != dev:
                gk = gk.to(device=dev, dtype=torch.float32)
            keep.append(gk)
            gptr.append(gk.data_ptr())
        same_dev = torch.cuda.current_device() == dev.index
        if gst is None:                                # rank-local: one call (coefficients + global + fine-grained backward)
            dv = torch.empty_like(v)
            dl = torch.empty_like(l)
            with (contextlib.nullcontext() if same_dev else torch.cuda.device(dev)):
                if ctx.peer is not None:
                    _lib.call("cfa_sparc_loss_gathered_bwd_ex", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code,
                              thr, scale, gw, lw, blk.data_ptr(), blk.numel(), *gptr, dv.data_ptr(), dl.data_ptr(), path,
                              ctx.peer[0], ctx.peer[1], ctx.gscale, _lib.stream_ptr())
                else:
                    _lib.call("cfa_sparc_loss_bwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr,
                              scale, gw, lw, blk.data_ptr(), blk.numel(), *gptr, dv.data_ptr(), dl.data_ptr(), path,
                              _lib.stream_ptr())
            return dv, dl, None, None, None, None, None, None, None, None, None, None
        with (contextlib.nullcontext() if same_dev else torch.cuda.device(dev)):
            coef = torch.empty(8, dtype=torch.float32, device=dev)
            _lib.call("cfa_sparc_coef_ptrs", *gptr, gw, lw, gst.Bg, ptr[1], coef.data_ptr(), _lib.stream_ptr())
            if ctx.gscale != 1.0:
                coef[:2] *= ctx.gscale                 # global term only (coef[2:4] are the rank-local fine-grained ones)
            dpv, dpl = _global_backward(gst, coef)
            dv = torch.empty_like(v)
            dl = torch.empty_like(l)
            sbytes = ctx.scratch.numel() if ctx.scratch is not None else 0
            sptr = ctx.scratch.data_ptr() if sbytes else 0
            _lib.call("cfa_sparc_bwd", v.data_ptr(), l.data_ptr(), mask_u8.data_ptr(), B, P, T, D, code, thr, scale,
                      ptr[5], ptr[2], ptr[3], ptr[6], ptr[7], gq[0], gq[1], coef.data_ptr() + 8, dpv.data_ptr(), dpl.data_ptr(),
                      dv.data_ptr(), dl.data_ptr(), sptr, sbytes, path, _lib.stream_ptr())
        return dv, dl, None, None, None, None, None, None, None, None, None, None


class _PairwiseFunction(torch.autograd.Function):
    """One-direction InfoNCE of a[B,D] vs b[B,D] (SPARCLoss.pairwise_contrastive_loss, losses.py:145-163)."""

    @staticmethod
    def forward(ctx, a, b, scale, eps):
        dev = _lib.require_cuda(a, b)
        a32 = a.detach().to(torch.float32).contiguous()
        b32 = b.detach().to(torch.float32).contiguous()
        with torch.cuda.device(dev):
            gst, sums = _global_forward(torch.stack([a32, b32]), scale, eps, 1, 0, None,
                                        path=1 if a.dtype == torch.float32 else 0)
        ctx.gst = gst
        ctx.dt = (a.dtype, b.dtype)
        return sums[0] / a.shape[0]

    @staticmethod
    def backward(ctx, g):
        gst = ctx.gst
        B = gst.a.shape[0]
        with torch.cuda.device(gst.a.device):
            z = torch.zeros(1, dtype=torch.float32, device=gst.a.device)
            coef = torch.cat([g.to(torch.float32).reshape(1) / B, z])      # only the a-rows direction carries loss
            da, db = _global_backward(gst, coef)
        return da.to(ctx.dt[0]), db.to(ctx.dt[1]), None, None


class _MaskedPairwiseFunction(torch.autograd.Function):
    """SPARCLoss.masked_pairwise_contrastive_loss(a[B,T,D], b[B,T,D], mask[B,T]) (losses.py:165-197)."""

    @staticmethod
    def forward(ctx, a, b, mask, scale):
        dev = _lib.require_cuda(a, b, mask)
        if a.dtype != b.dtype or a.dtype not in _lib.DTYPE_CODE:
            raise _lib.CfaError(f"masked_pairwise_contrastive_loss: a, b must share a dtype in fp32/bf16/fp16, got {a.dtype}, {b.dtype}")
        if a.dim() != 3 or a.shape != b.shape or mask.shape != a.shape[:2]:
            raise _lib.CfaError(f"masked_pairwise_contrastive_loss: bad shapes a{tuple(a.shape)} b{tuple(b.shape)} mask{tuple(mask.shape)}")
        a = a.contiguous()
        b = b.contiguous()
        mask_u8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else (mask != 0).contiguous().view(torch.uint8)
        B, T, D = a.shape
        buf = torch.empty(B * T + B + 2, dtype=torch.float32, device=dev)       # lse_row | partial | (loss, n_valid)
        p0 = buf.data_ptr()
        with torch.cuda.device(dev):
            _lib.call("cfa_masked_pairwise_fwd", a.data_ptr(), b.data_ptr(), mask_u8.data_ptr(), B, T, D,
                      _lib.DTYPE_CODE[a.dtype], scale, p0, p0 + 4 * B * T, p0 + 4 * (B * T + B), _lib.stream_ptr())
        ctx.save_for_backward(a, b, mask_u8, buf)
        ctx.scale = scale
        return buf[B * T + B].clone()

    @staticmethod
    def backward(ctx, g):
        a, b, mask_u8, buf = ctx.saved_tensors
        B, T, D = a.shape
        p0 = buf.data_ptr()
        g = g.to(device=a.device, dtype=torch.float32).contiguous()
        da = torch.empty_like(a)
        db = torch.empty_like(b)
        with torch.cuda.device(a.device):
            _lib.call("cfa_masked_pairwise_bwd", a.data_ptr(), b.data_ptr(), mask_u8.data_ptr(), B, T, D,
                      _lib.DTYPE_CODE[a.dtype], ctx.scale, p0, p0 + 4 * (B * T + B), g.data_ptr(), da.data_ptr(),
                      db.data_ptr(), _lib.stream_ptr())
        return da, db, None, None


class SPARCLoss(nn.Module):
    """SPARC loss (https://arxiv.org/abs/2401.09865), reference API: finetune/losses.py:136-264."""

    def __init__(self, config, gather=False, process_group=None, kernel_path: str = "auto",
                 fused_calls: bool = True, cast_to_bf16: bool = False, gather_grad_reduce: str = "mean"):
        super().__init__()
        # gather_grad_reduce (only with gather=True, world > 1): how the caller combines parameter gradients over ranks.
        #   "mean" (default, what DistributedDataParallel does, dist_finetuner.py:57): the gradient of the all-gathered
        #          GLOBAL term is multiplied by the world size, so that after DDP's average every parameter receives
        #          d(global mean loss)/d(theta) — all-gather-with-grad semantics, where the reduce-scatter sums the ranks'
        #          contributions before the DDP mean.  The rank-local fine-grained term needs no factor.
        #   "sum"  (or any scheme that adds the ranks' gradients): the plain d(global mean loss)/d(local rows), once per rank.
        if gather_grad_reduce not in ("mean", "sum"):
            raise ValueError(f"gather_grad_reduce must be 'mean' or 'sum', got {gather_grad_reduce!r}")
        self.gather_grad_reduce = gather_grad_reduce
        # cast_to_bf16 (opt-in): fp16 embeddings — what torch.autocast() hands the loss by default (finetuner.py:120) — and
        # fp32 embeddings (use_amp off, finetuner.py:156) are rounded to bf16 on entry so that the tcgen05 kernels run
        # (measured 17x faster at config 2 than the fp32-exact CUDA-core path those dtypes are otherwise routed to);
        # gradients come back in the caller's dtype.  This changes the inputs by up to 2^-9 relative (exactly what
        # autocast(dtype=torch.bfloat16) would have produced), so it is NOT within the parity bar and stays off by default.
        self.cast_to_bf16 = cast_to_bf16
        # gather: False = rank-local global loss (the reference under DDP); True = all-gathered over peer memory when the
        # ranks share an NVLink box and the shape allows, else over NCCL; "nccl" = always the NCCL all-gather path
        # fused_calls: rank-local loss through one library call per direction (cfa_sparc_loss_fwd / _bwd); False keeps
        # the per-stage entry points (same kernels, same numbers; used to time the stages separately)
        self.fused_calls = fused_calls
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
        """a, b: [B, T, D], mask: [B, T] -> masked token-level CE(sum over valid tokens) / (mask.sum() + 1e-8)."""
        if mask.dtype not in (torch.bool, torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            raise TypeError(f"mask must be bool or integer, got {mask.dtype}")
        return _MaskedPairwiseFunction.apply(a, b, mask, float(self.inverse_temperature))

    def forward(self, v_patch_embed: torch.Tensor, l_token_embed: torch.Tensor,
                language_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
        if language_mask.dtype not in (torch.bool, torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
            # the reference fails in `~language_mask` for float masks (SURVEY §8b errors)
            raise TypeError(f"language_mask must be bool or integer, got {language_mask.dtype}")
        if self.cast_to_bf16 and v_patch_embed.dtype in (torch.float16, torch.float32) \
                and l_token_embed.dtype == v_patch_embed.dtype:
            v_patch_embed = v_patch_embed.to(torch.bfloat16)       # autograd casts the bf16 gradients back
            l_token_embed = l_token_embed.to(torch.bfloat16)
        out = _SparcFunction.apply(v_patch_embed, l_token_embed, language_mask, float(self.similarity_threshold),
                                   float(self.global_loss_weight), float(self.local_loss_weight),
                                   float(self.inverse_temperature), self.gather, self.process_group, self.kernel_path,
                                   self.fused_calls, self.gather_grad_reduce == "mean")
        return dict(zip(SPARC_KEYS, out))                  # one autograd node, 7 outputs


# ----------------------------------------------------------------------------------------------
# CLIP InfoNCE
# ----------------------------------------------------------------------------------------------
class _ClipFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, txt, temperature, gather, group, ddp_mean=True):
        dev = _lib.require_cuda(img, txt)
        if img.dim() != 2 or img.shape != txt.shape:
            raise _lib.CfaError(f"CustomCLIPLoss: expected two [B,D] tensors, got {tuple(img.shape)}, {tuple(txt.shape)}")
        ab = torch.stack([img.detach().to(torch.float32), txt.detach().to(torch.float32)])
        ctx.peer = None
        with torch.cuda.device(dev):
            world, rank, group = _dist_ctx(group, gather)
            ctx.gscale = float(world) if (ddp_mean and world > 1) else 1.0      # see SPARCLoss.__init__ (gather_grad_reduce)
            B, D = img.shape
            if world > 1 and gather != "nccl" and img.dtype != torch.float32 \
                    and _L.cfa_global_infonce_path(B, world * B, D, 0) == 2:
                # all-gathered loss over peer memory (csrc/peer_exchange.cu): no collective call on the step path
                from . import peer as _peer
                ex = _peer.get_exchange(B, D, group)
                if ex is not None:
                    nbytes = _L.cfa_global_infonce_gathered_workspace_bytes(B, D, world)
                    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    sums = torch.empty(2, dtype=torch.float32, device=dev)
                    _lib.call("cfa_global_infonce_gathered_fwd", ab.data_ptr(), B, D, 1.0 / temperature, 0.0, ws.data_ptr(),
                              nbytes, world, rank, ex.blocks, ex.next_step(), sums.data_ptr(), _lib.stream_ptr())
                    ctx.peer = (world, rank, 1.0 / temperature)
                    ctx.save_for_backward(ab, ws)
                    ctx.dt = (img.dtype, txt.dtype)
                    return (sums[0] + sums[1]) * (0.5 / (world * B))
            # x / x.norm(): no eps in this loss (losses.py:17-18)
            gst, sums = _global_forward(ab, 1.0 / temperature, 0.0, world, rank, group,
                                        path=1 if img.dtype == torch.float32 else 0)
            loss = (sums[0] + sums[1]) * (0.5 / gst.Bg)            # mean CE both ways, averaged (losses.py:27-29)
        ctx.gst = gst
        ctx.dt = (img.dtype, txt.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.peer is not None:
            world, rank, scale = ctx.peer
            ab, ws = ctx.saved_tensors
            _, B, D = ab.shape
            with torch.cuda.device(ab.device):
                c = (g.to(torch.float32).reshape(1) * (ctx.gscale * 0.5 / (world * B))).expand(2).contiguous()
                dab = torch.empty_like(ab)
                _lib.call("cfa_global_infonce_gathered_bwd", ab.data_ptr(), B, D, scale, 0.0, ws.data_ptr(), ws.numel(),
                          c.data_ptr(), dab.data_ptr(), world, rank, _lib.stream_ptr())
            return dab[0].to(ctx.dt[0]), dab[1].to(ctx.dt[1]), None, None, None, None
        gst = ctx.gst
        with torch.cuda.device(gst.a.device):
            c = (g.to(torch.float32).reshape(1) * (ctx.gscale * 0.5 / gst.Bg)).expand(2).contiguous()
            da, db = _global_backward(gst, c)
        return da.to(ctx.dt[0]), db.to(ctx.dt[1]), None, None, None, None


class CustomCLIPLoss(nn.Module):
    """Symmetric CLIP InfoNCE, reference API: finetune/losses.py:7-36 (logits are DIVIDED by temperature)."""

    def __init__(self, temperature: float = 0.07, gather: bool = False, process_group=None,
                 gather_grad_reduce: str = "mean"):
        super().__init__()
        self.temperature = temperature
        self.gather = gather
        self.process_group = process_group
        if gather_grad_reduce not in ("mean", "sum"):
            raise ValueError(f"gather_grad_reduce must be 'mean' or 'sum', got {gather_grad_reduce!r}")
        self.gather_grad_reduce = gather_grad_reduce       # see SPARCLoss.__init__

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor,
                custom_features: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        clip_loss = _ClipFunction.apply(image_features, text_features, float(self.temperature), self.gather,
                                        self.process_group, self.gather_grad_reduce == "mean")
        return {"clip_loss": clip_loss, "total_loss": clip_loss}


# ----------------------------------------------------------------------------------------------
# Counting losses: CLIPCountLoss (losses.py:39-133), CountLoss (losses.py:267-309)
# ----------------------------------------------------------------------------------------------
class _LogitsCEFunction(torch.autograd.Function):
    """(CE(img_logits, arange) + CE(text_logits, arange)) / 2 on caller-provided [B,B] logits (losses.py:276-279)."""

    @staticmethod
    def forward(ctx, la, lb):
        dev = _lib.require_cuda(la, lb)
        if la.dim() != 2 or la.shape[0] != la.shape[1] or lb.shape != la.shape or la.dtype != lb.dtype \
                or la.dtype not in _lib.DTYPE_CODE:
            raise _lib.CfaError(f"CountLoss: expected two [B,B] logits of one dtype, got {tuple(la.shape)}, {tuple(lb.shape)}")
        la = la.contiguous(); lb = lb.contiguous()
        B = la.shape[0]
        buf = torch.empty(4 * B + 1, dtype=torch.float32, device=dev)          # lse2 [2,B] | ce2 [2,B] | out
        p0 = buf.data_ptr()
        with torch.cuda.device(dev):
            _lib.call("cfa_logits_ce_fwd", la.data_ptr(), lb.data_ptr(), B, _lib.DTYPE_CODE[la.dtype], p0, p0 + 8 * B,
                      p0 + 16 * B, _lib.stream_ptr())
        ctx.save_for_backward(la, lb, buf)
        return buf[4 * B].clone()

    @staticmethod
    def backward(ctx, g):
        la, lb, buf = ctx.saved_tensors
        B = la.shape[0]
        g = g.to(device=la.device, dtype=torch.float32).contiguous()
        da, db = torch.empty_like(la), torch.empty_like(lb)
        with torch.cuda.device(la.device):
            _lib.call("cfa_logits_ce_bwd", la.data_ptr(), lb.data_ptr(), B, _lib.DTYPE_CODE[la.dtype], buf.data_ptr(),
                      g.data_ptr(), da.data_ptr(), db.data_ptr(), _lib.stream_ptr())
        return da, db


class _CountContrastiveFunction(torch.autograd.Function):
    """mean_b [ log sum_c exp(e_i . e_cf[c] / T) - e_i . e_k / T ] on L2-normalised rows (losses.py:281-301)."""

    @staticmethod
    def forward(ctx, ei, ek, ek_cf, temperature, include_pos):
        dev = _lib.require_cuda(ei, ek, ek_cf)
        if ei.dim() != 2 or ek.shape != ei.shape or ek_cf.dim() != 3 or ek_cf.shape[0] != ei.shape[0] \
                or ek_cf.shape[2] != ei.shape[1]:
            raise _lib.CfaError(f"CountLoss: bad shapes ei{tuple(ei.shape)} ek{tuple(ek.shape)} ek_cf{tuple(ek_cf.shape)}")
        if not (ei.dtype == ek.dtype == ek_cf.dtype) or ei.dtype not in _lib.DTYPE_CODE:
            raise _lib.CfaError("CountLoss: embeddings must share a dtype in fp32/bf16/fp16")
        ei = ei.contiguous(); ek = ek.contiguous(); ek_cf = ek_cf.contiguous()
        B, D = ei.shape
        C = ek_cf.shape[1]
        buf = torch.empty(B + 1, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("cfa_count_contrastive_fwd", ei.data_ptr(), ek.data_ptr(), ek_cf.data_ptr(), B, C, D,
                      _lib.DTYPE_CODE[ei.dtype], temperature, int(include_pos), buf.data_ptr(), buf.data_ptr() + 4 * B,
                      _lib.stream_ptr())
        ctx.save_for_backward(ei, ek, ek_cf)
        ctx.hp = (temperature, int(include_pos))
        return buf[B].clone()

    @staticmethod
    def backward(ctx, g):
        ei, ek, ek_cf = ctx.saved_tensors
        B, D = ei.shape
        C = ek_cf.shape[1]
        g = g.to(device=ei.device, dtype=torch.float32).contiguous()
        dei, dek, dcf = torch.empty_like(ei), torch.empty_like(ek), torch.empty_like(ek_cf)
        with torch.cuda.device(ei.device):
            _lib.call("cfa_count_contrastive_bwd", ei.data_ptr(), ek.data_ptr(), ek_cf.data_ptr(), B, C, D,
                      _lib.DTYPE_CODE[ei.dtype], ctx.hp[0], ctx.hp[1], g.data_ptr(), dei.data_ptr(), dek.data_ptr(),
                      dcf.data_ptr(), _lib.stream_ptr())
        return dei, dek, dcf, None, None


class CountLoss(nn.Module):
    """Reference API: finetune/losses.py:267-309 (used by count_finetuner.py:125-131).
    forward(img_logits[B,B], text_logits[B,B], ei[B,D], ek[B,D], ek_cf[B,C,D]) -> clip_loss, count_loss, total_loss."""

    def __init__(self, temperature: float = 0.07, alpha=1.0):
        super().__init__()
        self.temperature = temperature
        self.alpha = alpha

    def forward(self, img_logits, text_logits, ei, ek, ek_cf) -> Dict[str, torch.Tensor]:
        clip_loss = _LogitsCEFunction.apply(img_logits, text_logits)
        count_loss = _CountContrastiveFunction.apply(ei, ek, ek_cf, float(self.temperature), False)
        total_loss = clip_loss + self.alpha * count_loss                          # losses.py:303
        return {"clip_loss": clip_loss, "count_loss": count_loss, "total_loss": total_loss}


class CLIPCountLoss(nn.Module):
    """Reference API: finetune/losses.py:39-133.  image_features [B,D], text_features [E,D] with E = B * templates:
    the image rows are repeated per template and the CLIP InfoNCE runs on the expanded E x E problem (global InfoNCE
    kernels).  The reference's count term groups `counts.size(0) // E` captions per expanded row out of the SAME [E,D]
    text matrix (:51,:69-86): with one count per caption the group is the positive alone and the term is exactly 0 (its
    gradient too); any other group size indexes past the text matrix and raises IndexError in the reference — kept."""

    def __init__(self, temperature: float = 0.07, count_alpha: float = 0.5, gather: bool = False, process_group=None):
        super().__init__()
        self.temperature = temperature
        self.count_alpha = count_alpha
        self.gather = gather
        self.process_group = process_group

    def count_loss(self, ei: torch.Tensor, ek: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        batch_size = ei.size(0)
        group_size = counts.size(0) // batch_size
        if group_size != 1:
            raise IndexError("index 0 is out of bounds for dimension 0 with size 0")     # what losses.py:76 ends in
        return torch.zeros((), dtype=torch.float64, device=ei.device)                    # -log(num / (num + 0)), fp64 (:55)

    def forward(self, image_features: torch.Tensor, text_features: torch.Tensor,
                count_features: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        batch_size = image_features.size(0)
        expanded = text_features.size(0)
        num_templates = expanded // batch_size                                            # losses.py:100-101
        image_exp = image_features.repeat_interleave(num_templates, dim=0)                # losses.py:104
        clip_loss = _ClipFunction.apply(image_exp, text_features, float(self.temperature), self.gather, self.process_group)
        count_loss = torch.tensor(0.0, device=image_features.device)                     # losses.py:117
        if count_features is not None:
            count_loss = self.count_loss(image_exp, text_features, count_features) * self.count_alpha
        total_loss = clip_loss + count_loss
        return {"clip_loss": clip_loss, "count_loss": count_loss, "total_loss": total_loss}
