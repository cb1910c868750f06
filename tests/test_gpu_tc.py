"""tcgen05 / TMA building blocks: every shared-memory operand flavour used by the tensor-core kernels,
checked against torch on exactly representable bf16 inputs (fp32 accumulation -> tight tolerance)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(a_mode, b_mode, N, K):
    from clip_finegrained_alignment_b200 import _lib
    g = torch.Generator().manual_seed(100 * a_mode + 10 * b_mode + N + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16)          # logical A[m, k]
    B = torch.randn(N, K, generator=g).to(torch.bfloat16)            # logical B[n, k]
    ref = A.float() @ B.float().t()
    a_src = (A.t().contiguous() if a_mode in (2, 3) else A).cuda()
    b_src = (B.t().contiguous() if b_mode in (1, 3, 5) else B).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("cfa_tc_selftest", a_mode, b_mode, N, K, a_src.data_ptr(), b_src.data_ptr(), D.data_ptr(),
              _lib.stream_ptr())
    torch.cuda.synchronize()
    return D.cpu(), ref


@pytest.mark.parametrize("a_mode,b_mode,N,K", [
    (0, 0, 208, 128),    # TMA SW128 K-major x TMA SW128 K-major        (S = l . v^T)
    (0, 0, 80, 64),
    (1, 0, 208, 128),    # interleaved K-major A x TMA K-major B          (L += G_kb . l_kb^T, dW += dG_kb . v_kb^T)
    (1, 1, 64, 208),     # interleaved K-major A x TMA tile read MN-major (G_kb = W . v_kb, X = dL . l_kb)
    (1, 1, 64, 80),
    (2, 1, 64, 80),      # interleaved read MN-major A x TMA MN-major B   (dv_kb = dS^T . l_kb)
    (2, 3, 64, 80),      # both interleaved, both MN-major                (dv_kb += W^T . dG_kb)
    (1, 2, 96, 64),
    (1, 3, 64, 208),     # interleaved K-major A x interleaved MN-major B (dl_kb += dL^T . G_kb uses a_mode 2)
    (4, 4, 208, 64),     # 32-wide blocks, SWIZZLE_64B tiles (backward kernel): S
    (4, 4, 80, 96),
    (1, 4, 208, 32),     # dW += dG_kb . v_kb^T
    (1, 5, 32, 208),     # G_kb = W . v_kb ; dl_kb = dS . v_kb
    (1, 5, 32, 80),      # X_kb = dL . l_kb
    (2, 5, 32, 80),      # dv_kb = dS^T . l_kb
    (2, 3, 32, 80),      # dv_kb += W^T . dG_kb ; dl_kb += dL^T . G_kb
    # third-generation ("transposed") kernels: raw tiles are the A operand, on-chip hi|lo operands are stacked along N
    (3, 3, 160, 208),    # TMA tile pair read MN-major (M = 128 d over two swizzle atoms) x interleaved MN-major (G'^T = v^T . Theta^T)
    (3, 3, 160, 96),
    (3, 2, 208, 80),     # l^T / G^T (MN-major TMA) x interleaved K-major, N = p                      (dv^T)
    (2, 3, 160, 208),    # interleaved MN-major A (M = token) x interleaved MN-major B                 (L' = Theta . S_raw^T)
    (0, 0, 240, 128),    # v tile x stacked [l; G hi; G lo] tiles                                      (backward P1)
    (1, 2, 80, 80),      # S_raw^T (interleaved K-major) x dLhat (interleaved K-major)                 (dW^T)
    (1, 3, 80, 80),      # W^T (interleaved K-major) x dLhat (interleaved MN-major)                    (Z^T)
])
def test_tcgen05_operand_modes(a_mode, b_mode, N, K):
    D, ref = _run(a_mode, b_mode, N, K)
    err = (D - ref).abs().max().item()
    assert torch.isfinite(D).all(), (a_mode, b_mode)
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), (a_mode, b_mode, N, K, err)


# ----------------------------------------------------------------------------------------------
# tensor-core SPARC kernels vs the oracle (oracle-A: fp64 reference math on the bf16-representable inputs)
# ----------------------------------------------------------------------------------------------
import types

from conftest import rel_err
from oracle import losses_oracle as lo


def _cfg(thr, gw=1.0, lw=1.0, s=1.0):
    return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=gw, local_loss_weight=lw,
                                 inverse_temperature=s)


def _sparc_tc(v, l, m, c, key="total_loss", path="tc"):
    from clip_finegrained_alignment_b200 import SPARCLoss
    vv = v.cuda().requires_grad_(True)
    ll = l.cuda().requires_grad_(True)
    out = SPARCLoss(c, kernel_path=path)(vv, ll, m.cuda())
    out[key].backward()
    torch.cuda.synchronize()
    return {k: x.detach().float().cpu() for k, x in out.items()}, vv.grad, ll.grad


@pytest.mark.parametrize("B,P,T,D,s", [(3, 196, 77, 512, 1.0), (2, 197, 77, 512, 2.0), (4, 50, 77, 256, 1.0),
                                       (2, 33, 20, 256, 3.0), (2, 256, 77, 768, 1.0), (5, 64, 128, 256, 1.0)])
def test_sparc_tc_vs_oracle(B, P, T, D, s):
    g = torch.Generator().manual_seed(B * 7 + P)
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool)
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    out, dv, dl = _sparc_tc(v, l, m, _cfg(thr, 0.9, 1.1, s))
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 0.9, 1.1, s)
    rv, rl = lo.sparc_backward(o)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), (k, float(out[k]), float(o[k]))
    # gradients come back in bf16: allow one output rounding (2^-9 relative, rms) on top of rtol 1e-3
    assert rel_err(dv.float(), rv) <= 1e-3 + 2.0 ** -8
    assert rel_err(dl.float(), rl) <= 1e-3 + 2.0 ** -8


@pytest.mark.parametrize("B,P,T,D,s,dtype", [
    (2, 196, 77, 512, 50.0, torch.bfloat16),     # |scale| > 30: max-subtracted log-sum-exp branch of sparc_fwd3 (CLIP logit scales reach 100)
    (2, 196, 77, 512, 100.0, torch.float16),
    (3, 1, 1, 128, 1.0, torch.bfloat16),         # one patch, one token
    (2, 17, 3, 128, 2.0, torch.bfloat16),        # tiny ragged tile
    (2, 255, 80, 256, 1.0, torch.bfloat16),      # T = NT = 80: no spare column (pooled image mean from the CUDA-core side job)
    (2, 130, 33, 640, 1.0, torch.float16),       # D = 5 x 128, second patch block almost empty
])
def test_sparc_gen3_edge_shapes_vs_oracle(B, P, T, D, s, dtype):
    import os
    if os.environ.get("CFA_SPARC_GEN") == "2":
        pytest.skip("third-generation kernels switched off (A/B aid)")
    from clip_finegrained_alignment_b200 import _lib
    code = _lib.DTYPE_CODE[dtype]
    assert _lib.lib.cfa_sparc_path(P, T, D, code, 0) == 2
    g = torch.Generator().manual_seed(P * 13 + T)
    v = torch.randn(B, P, D, generator=g).to(dtype)
    l = torch.randn(B, T, D, generator=g).to(dtype)
    m = torch.ones(B, T, dtype=torch.bool)
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    out, dv, dl = _sparc_tc(v, l, m, _cfg(thr, 0.9, 1.1, s))
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 0.9, 1.1, s)
    rv, rl = lo.sparc_backward(o)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), (k, float(out[k]), float(o[k]))
    assert torch.isfinite(dv).all() and torch.isfinite(dl).all()
    if float(rv.norm()) > 0:
        assert rel_err(dv.float(), rv) <= 1e-3 + 2.0 ** -8, rel_err(dv.float(), rv)
    if float(rl.norm()) > 0:
        assert rel_err(dl.float(), rl) <= 1e-3 + 2.0 ** -8, rel_err(dl.float(), rl)


def test_sparc_gen3_fully_masked_sample():
    """One sample without a single valid token ("truncate" semantics: it contributes nothing to the local loss; the
    clamped token count keeps its pooled text mean at zero): finite losses and gradients, equal to the oracle."""
    B, P, T, D = 3, 196, 77, 512
    g = torch.Generator().manual_seed(77)
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool); m[1, :] = False; m[2, 5:] = False
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    out, dv, dl = _sparc_tc(v, l, m, _cfg(thr))
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 1.0, 1.0, 1.0, mask_semantics="truncate")
    rv, rl = lo.sparc_backward(o)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), (k, float(out[k]), float(o[k]))
    assert torch.isfinite(dv).all() and torch.isfinite(dl).all()
    assert rel_err(dv.float(), rv) <= 1e-3 + 2.0 ** -8 and rel_err(dl.float(), rl) <= 1e-3 + 2.0 ** -8
    assert float(dl[1].abs().max()) == 0.0


@pytest.mark.parametrize("B,P,T,D,s,mag,up", [
    (3, 197, 77, 512, 14.0, 0.5, 65536.0),      # GradScaler-scaled upstream gradient (finetuner.py:120-134)
    (2, 196, 77, 512, 1.0, 1.0, 1.0),           # no loss scale: coefficients ~ 1e-5
    (2, 196, 77, 512, 1.0, 0.03, 1.0),          # small embeddings: S_raw ~ 1e-2
    (2, 196, 77, 512, 5.0, 6.0, 1024.0),        # large embeddings: |v . l| ~ 1e3
    (4, 50, 40, 256, 1.0, 1.0, 1.0),            # generic (non-flagship) instantiation, padded mask below
])
def test_sparc_fp16_on_tensor_cores(B, P, T, D, s, mag, up):
    """fp16 embeddings (torch.autocast's default; finetuner.py:51,120-134 with GradScaler) on the tcgen05 path: tcgen05
    kind::f16 takes no mixed fp16 x bf16 operand pair, so the on-chip operands are fp16 hi|lo, kept in range by two
    per-sample powers of two (csrc/sparc_tc_bwd3.cu).  fp64 oracle on the fp16-representable inputs, 1e-3."""
    import os
    if os.environ.get("CFA_SPARC_GEN") == "2":
        pytest.skip("fp16 needs the third-generation kernels (switched off: A/B aid)")
    from clip_finegrained_alignment_b200 import SPARCLoss, _lib
    g = torch.Generator().manual_seed(B * 31 + P)
    v0 = (torch.randn(B, P, D, generator=g) * mag).to(torch.float16)
    l0 = (torch.randn(B, T, D, generator=g) * mag).to(torch.float16)
    m = torch.ones(B, T, dtype=torch.bool)
    if T == 40:
        m[0, 25:] = False; m[2, 1:] = False
    assert _lib.lib.cfa_sparc_path(P, T, D, _lib.DTYPE_CODE[torch.float16], 0) == 2
    assert _lib.lib.cfa_sparc_bwd_path(P, T, D, _lib.DTYPE_CODE[torch.float16], 0) == 2
    v = v0.cuda().requires_grad_(True); l = l0.cuda().requires_grad_(True)
    out = SPARCLoss(_cfg(1.0 / P, 1.0, 1.0, s), kernel_path="tc")(v, l, m.cuda())
    (out["total_loss"] * up).backward()
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    # padded masks: the documented "truncate" semantics (the reference itself returns NaN for any False entry)
    o = lo.sparc_forward(v0.double(), l0.double(), m, thr, 1.0, 1.0, s, mask_semantics="truncate" if T == 40 else "reference")
    rv, rl = lo.sparc_backward(o, {"total_loss": up})
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), (k, float(out[k]), float(o[k]))
    assert v.grad.dtype == torch.float16 and torch.isfinite(v.grad).all() and torch.isfinite(l.grad).all()
    # gradients come back in fp16 (2^-11 relative rounding, plus subnormal flushing of the smallest entries when the
    # upstream gradient is not loss-scaled -- exactly what torch's own fp16 backward produces)
    assert rel_err(v.grad.float(), rv) <= 1e-3, rel_err(v.grad.float(), rv)
    assert rel_err(l.grad.float(), rl) <= 1e-3, rel_err(l.grad.float(), rl)


def test_sparc_fp16_cuda_core_path_still_exact():
    """kernel_path="simt" keeps the fp32-exact CUDA-core kernels for fp16 inputs."""
    from clip_finegrained_alignment_b200 import SPARCLoss
    B, P, T, D, s = 2, 197, 77, 512, 14.0
    g = torch.Generator().manual_seed(5)
    v0 = (torch.randn(B, P, D, generator=g) * 0.5).to(torch.float16)
    l0 = (torch.randn(B, T, D, generator=g) * 0.5).to(torch.float16)
    m = torch.ones(B, T, dtype=torch.bool)
    v = v0.cuda().requires_grad_(True); l = l0.cuda().requires_grad_(True)
    out = SPARCLoss(_cfg(1.0 / P, 1.0, 1.0, s), kernel_path="simt")(v, l, m.cuda())
    (out["total_loss"] * 65536.0).backward()
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    o = lo.sparc_forward(v0.double(), l0.double(), m, thr, 1.0, 1.0, s)
    rv, rl = lo.sparc_backward(o, {"total_loss": 65536.0})
    assert rel_err(v.grad.float(), rv) <= 1e-3 and rel_err(l.grad.float(), rl) <= 1e-3


def test_sparc_tc_padded_mask():
    g = torch.Generator().manual_seed(3)
    B, P, T, D = 4, 196, 77, 256
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    lens = torch.tensor([77, 5, 40, 76])
    m = torch.arange(T)[None, :] < lens[:, None]
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    out, dv, dl = _sparc_tc(v, l, m, _cfg(thr))
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 1.0, 1.0, 1.0, mask_semantics="truncate")
    rv, rl = lo.sparc_backward(o)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), k
    assert rel_err(dv.float(), rv) <= 1e-3 + 2.0 ** -8
    assert rel_err(dl.float(), rl) <= 1e-3 + 2.0 ** -8


def test_sparc_tc_matches_simt_path():
    g = torch.Generator().manual_seed(11)
    B, P, T, D = 6, 196, 77, 512
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool)
    a, dva, dla = _sparc_tc(v, l, m, _cfg(1.0 / P), path="tc")
    b, dvb, dlb = _sparc_tc(v, l, m, _cfg(1.0 / P), path="simt")
    for k in a:
        assert abs(float(a[k]) - float(b[k])) <= 2e-5 * max(1.0, abs(float(b[k]))), k
    assert rel_err(dva.float(), dvb.float()) <= 5e-3
    assert rel_err(dla.float(), dlb.float()) <= 5e-3


@pytest.mark.parametrize("dtype,P,D", [(torch.bfloat16, 196, 512), (torch.float32, 50, 64), (torch.bfloat16, 576, 768)])
def test_fused_calls_equal_staged_calls(dtype, P, D):
    """cfa_sparc_loss_fwd / _bwd (one library call per direction) run the same kernels as the per-stage entry points:
    bit-identical losses and gradients, for the tensor-core, the CUDA-core and the global-scratch (config 4) paths."""
    from clip_finegrained_alignment_b200 import SPARCLoss
    g = torch.Generator().manual_seed(3)
    B, T = 5, 77
    v0 = torch.randn(B, P, D, generator=g).to(dtype); l0 = torch.randn(B, T, D, generator=g).to(dtype)
    m = torch.ones(B, T, dtype=torch.bool); m[1, 40:] = False
    res = []
    for fused in (True, False):
        v = v0.cuda().requires_grad_(True); l = l0.cuda().requires_grad_(True)
        out = SPARCLoss(_cfg(1.0 / P, 0.7, 1.3, 2.0), fused_calls=fused)(v, l, m.cuda())
        (out["total_loss"] * 3.0 + out["loss_vl_local"]).backward()
        res.append(({k: float(x) for k, x in out.items()}, v.grad.clone(), l.grad.clone()))
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_sparc_tc_bitwise_deterministic():
    """No atomics with run-dependent ordering anywhere on the tensor-core path: repeated evaluations of the same step
    (32 samples -> many CTAs in flight) return identical bits.  Catches races as well."""
    from clip_finegrained_alignment_b200 import SPARCLoss
    g = torch.Generator().manual_seed(21)
    B, P, T, D = 32, 196, 77, 512
    v0 = torch.randn(B, P, D, generator=g).to(torch.bfloat16).cuda(); l0 = torch.randn(B, T, D, generator=g).to(torch.bfloat16).cuda()
    m = torch.ones(B, T, dtype=torch.bool, device="cuda")
    crit = SPARCLoss(_cfg(1.0 / P))
    ref = None
    for _ in range(6):
        v = v0.clone().requires_grad_(True); l = l0.clone().requires_grad_(True)
        out = crit(v, l, m)
        out["total_loss"].backward()
        cur = (torch.stack([out[k].detach() for k in lo.SPARC_KEYS]), v.grad, l.grad)
        if ref is None:
            ref = cur
        else:
            assert all(torch.equal(a, b) for a, b in zip(ref, cur))


# ----------------------------------------------------------------------------------------------
# tensor-core global InfoNCE (tcgen05 logits tiles, bf16 hi/lo-split normalised operands)
# ----------------------------------------------------------------------------------------------
# (19, 128, 64): 19 column tiles per row -> 38 softmax partials (> 32: global_merge_rows_kernel) and the red.global dA path
@pytest.mark.parametrize("N,B,D,s", [(1, 256, 512, 1.0), (1, 100, 64, 14.285), (3, 70, 256, 5.0), (4, 128, 512, 2.0),
                                     (19, 128, 64, 3.0)])
def test_global_infonce_tc_emulated_ranks(N, B, D, s):
    """path=2 (tensor cores) through the C ABI: N emulated ranks of B rows against N*B gathered columns, both
    directions, forward (lse, CE sums) and backward, vs the fp64 oracle on the concatenated batch."""
    from clip_finegrained_alignment_b200 import _lib
    g = torch.Generator().manual_seed(N * 1000 + B + D)
    Bg = N * B
    a = torch.randn(Bg, D, generator=g)
    b = torch.randn(Bg, D, generator=g)
    f1 = lo.infonce_forward(a.double(), b.double(), s)
    f2 = lo.infonce_forward(b.double(), a.double(), s)
    da_ref, db_ref = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"], s,
                                                   0.5, 0.5, float(Bg))
    ac, bc = a.cuda(), b.cuda()
    assert _lib.lib.cfa_global_infonce_path(B, Bg, D, 2) == 2
    ws_bytes = _lib.lib.cfa_global_infonce_workspace_bytes(B, Bg, D)
    wss, lse, norms = [], [], []
    sums = torch.zeros(2, device="cuda")
    for r in range(N):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        l2 = torch.empty(2, B, device="cuda"); n2 = torch.empty(2, B, device="cuda"); s2 = torch.empty(2, device="cuda")
        al, bl = ac[r * B:(r + 1) * B].contiguous(), bc[r * B:(r + 1) * B].contiguous()
        _lib.call("cfa_global_infonce_fwd", al.data_ptr(), bl.data_ptr(), ac.data_ptr(), bc.data_ptr(), B, Bg, D, r * B, s,
                  1e-12, l2.data_ptr(), n2.data_ptr(), s2.data_ptr(), 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), ws_bytes, 2,
                  0, _lib.stream_ptr())
        wss.append(ws); lse.append(l2); norms.append(n2); sums += s2
    ref_loss = float(0.5 * (f1["loss_sum"] + f2["loss_sum"]) / Bg)
    assert abs(float(0.5 * sums.sum() / Bg) - ref_loss) <= 2e-5 * max(1.0, ref_loss)
    lse_all = torch.cat(lse, dim=1).contiguous()
    torch.testing.assert_close(lse_all[0].cpu().double(), f1["lse"], rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(lse_all[1].cpu().double(), f2["lse"], rtol=2e-5, atol=2e-5)
    coef = torch.full((2,), 0.5 / Bg, device="cuda")
    for r in range(N):
        sl = slice(r * B, (r + 1) * B)
        al, bl = ac[sl].contiguous(), bc[sl].contiguous()
        da = torch.empty(B, D, device="cuda"); db = torch.empty(B, D, device="cuda")
        _lib.call("cfa_global_infonce_bwd", al.data_ptr(), bl.data_ptr(), ac.data_ptr(), bc.data_ptr(), B, Bg, D, r * B, s,
                  1e-12, lse[r].data_ptr(), lse_all.data_ptr(), norms[r].data_ptr(), coef.data_ptr(), da.data_ptr(),
                  db.data_ptr(), wss[r].data_ptr(), ws_bytes, 2, 0, _lib.stream_ptr())
        torch.cuda.synchronize()
        assert rel_err(da, da_ref[sl]) <= 2e-4, (r, rel_err(da, da_ref[sl]))
        assert rel_err(db, db_ref[sl]) <= 2e-4, (r, rel_err(db, db_ref[sl]))


def test_global_infonce_tc_raw_allgather_layout():
    """gathered_ranks > 1: the kernels consume the RAW all-gather outputs ([ranks][2][B][D] embeddings and
    [ranks][2B+2] lse/sum packs) — same numbers as the re-laid-out path, no copy kernels in between."""
    from clip_finegrained_alignment_b200 import _lib
    N, B, D, s = 3, 96, 256, 4.0
    Bg = N * B
    g = torch.Generator().manual_seed(77)
    a = torch.randn(Bg, D, generator=g); b = torch.randn(Bg, D, generator=g)
    f1 = lo.infonce_forward(a.double(), b.double(), s)
    f2 = lo.infonce_forward(b.double(), a.double(), s)
    da_ref, db_ref = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"], s,
                                                   0.5, 0.5, float(Bg))
    gathered = torch.stack([torch.stack([a[r * B:(r + 1) * B], b[r * B:(r + 1) * B]]) for r in range(N)]).cuda()   # [N,2,B,D]
    packs = torch.zeros(N, 2 * B + 2, device="cuda")
    ws_bytes = _lib.lib.cfa_global_infonce_workspace_bytes(B, Bg, D)
    wss, norms = [], []
    for r in range(N):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        n2 = torch.empty(2, B, device="cuda")
        loc = gathered[r]
        _lib.call("cfa_global_infonce_fwd", loc[0].data_ptr(), loc[1].data_ptr(), gathered[0, 0].data_ptr(),
                  gathered[0, 1].data_ptr(), B, Bg, D, r * B, s, 1e-12, packs[r].data_ptr(), n2.data_ptr(),
                  packs[r].data_ptr() + 8 * B, 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), ws_bytes, 2, N, _lib.stream_ptr())
        wss.append(ws); norms.append(n2)
    torch.cuda.synchronize()
    ref_loss = float(0.5 * (f1["loss_sum"] + f2["loss_sum"]) / Bg)
    assert abs(float(0.5 * packs[:, 2 * B:].sum() / Bg) - ref_loss) <= 2e-5 * max(1.0, ref_loss)
    coef = torch.full((2,), 0.5 / Bg, device="cuda")
    for r in range(N):
        loc = gathered[r]
        da = torch.empty(B, D, device="cuda"); db = torch.empty(B, D, device="cuda")
        _lib.call("cfa_global_infonce_bwd", loc[0].data_ptr(), loc[1].data_ptr(), gathered[0, 0].data_ptr(),
                  gathered[0, 1].data_ptr(), B, Bg, D, r * B, s, 1e-12, packs[r].data_ptr(), packs.data_ptr(),
                  norms[r].data_ptr(), coef.data_ptr(), da.data_ptr(), db.data_ptr(), wss[r].data_ptr(), ws_bytes, 2, N,
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        sl = slice(r * B, (r + 1) * B)
        assert rel_err(da, da_ref[sl]) <= 2e-4 and rel_err(db, db_ref[sl]) <= 2e-4
    # scalar epilogue straight from the gathered packs
    out8 = torch.zeros(8, device="cuda")
    part = torch.zeros(B, 2, device="cuda"); mask = torch.ones(B, 5, dtype=torch.uint8, device="cuda")
    _lib.call("cfa_sparc_finalize", packs.data_ptr(), Bg, part.data_ptr(), mask.data_ptr(), B, 5, 1.0, 1.0, out8.data_ptr(), N,
              _lib.stream_ptr())
    assert abs(float(out8[0]) - ref_loss) <= 2e-5 * max(1.0, ref_loss)


def test_clip_loss_bf16_uses_tensor_cores():
    from clip_finegrained_alignment_b200 import CustomCLIPLoss
    g = torch.Generator().manual_seed(2)
    B, D = 300, 512
    a0 = torch.randn(B, D, generator=g).to(torch.bfloat16)
    b0 = torch.randn(B, D, generator=g).to(torch.bfloat16)
    a = a0.cuda().requires_grad_(True); b = b0.cuda().requires_grad_(True)
    out = CustomCLIPLoss(0.07)(a, b)
    out["total_loss"].backward()
    o = lo.clip_loss_forward(a0.double(), b0.double(), 0.07)
    da, db = lo.clip_loss_backward(o, 0.07)
    assert abs(float(out["clip_loss"]) - float(o["clip_loss"])) <= 1e-4 * float(o["clip_loss"])
    assert rel_err(a.grad.float(), da) <= 1e-3 + 2.0 ** -8
    assert rel_err(b.grad.float(), db) <= 1e-3 + 2.0 ** -8


def test_sparc_cast_to_bf16_opt_in():
    """cast_to_bf16=True: fp16 (or fp32) inputs are rounded to bf16 and take the tensor-core kernels; the result equals the bf16
    path on the rounded inputs exactly, and gradients come back in fp16."""
    from clip_finegrained_alignment_b200 import SPARCLoss
    g = torch.Generator().manual_seed(9)
    B, P, T, D = 3, 50, 77, 256
    v16 = torch.randn(B, P, D, generator=g).half().cuda()
    l16 = torch.randn(B, T, D, generator=g).half().cuda()
    m = torch.ones(B, T, dtype=torch.bool, device="cuda")
    c = _cfg(1.0 / P)
    va = v16.clone().requires_grad_(True); la = l16.clone().requires_grad_(True)
    oa = SPARCLoss(c, cast_to_bf16=True)(va, la, m)
    oa["total_loss"].backward()
    vb = v16.to(torch.bfloat16).requires_grad_(True); lb = l16.to(torch.bfloat16).requires_grad_(True)
    ob = SPARCLoss(c)(vb, lb, m)
    ob["total_loss"].backward()
    assert va.grad.dtype == torch.float16 and la.grad.dtype == torch.float16
    assert float(oa["total_loss"]) == float(ob["total_loss"])
    assert torch.equal(va.grad.float(), vb.grad.float().half().float()) and torch.equal(la.grad.float(), lb.grad.float().half().float())


@pytest.mark.parametrize("B,P,T,D,dtype", [(5, 196, 77, 512, torch.bfloat16), (3, 197, 77, 512, torch.float16),
                                           (4, 50, 20, 128, torch.bfloat16), (2, 255, 80, 256, torch.bfloat16)])
def test_sparc_tc_stays_inside_its_buffers(B, P, T, D, dtype):
    """Guard bands (compute-sanitizer is not available on this pool): the inputs, the workspace and the two gradient outputs
    of the one-call C ABI sit between canary regions inside larger allocations; no kernel of the forward / backward chain
    (TMA loads with out-of-range rows, 16-byte transposed stores, the saved G planes, statistics, partial buffers) may
    write a byte outside what it was given, and the inputs must come back unchanged."""
    from clip_finegrained_alignment_b200 import _lib
    L = _lib.lib
    code = _lib.DTYPE_CODE[dtype]
    guard = 4096                                          # bytes on each side, 128-byte aligned payloads
    def banded(nbytes):
        n = (nbytes + 127) & ~127
        buf = torch.full((guard + n + guard,), 0xA5, dtype=torch.uint8, device="cuda")
        return buf, buf[guard:guard + nbytes]
    g = torch.Generator().manual_seed(B + P)
    esz = 2
    vb, vpay = banded(B * P * D * esz); lb, lpay = banded(B * T * D * esz)
    v = vpay.view(dtype).view(B, P, D); l = lpay.view(dtype).view(B, T, D)
    v.copy_(torch.randn(B, P, D, generator=g).to(dtype)); l.copy_(torch.randn(B, T, D, generator=g).to(dtype))
    v0, l0 = v.clone(), l.clone()
    mask = torch.ones(B, T, dtype=torch.uint8, device="cuda"); mask[0, T // 2:] = 0
    nws = L.cfa_sparc_loss_workspace_bytes(B, P, T, D, code, 0)
    wb, wpay = banded(nws)
    db, dpay = banded(B * P * D * esz); eb, epay = banded(B * T * D * esz)
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    _lib.call("cfa_sparc_loss_fwd", v.data_ptr(), l.data_ptr(), mask.data_ptr(), B, P, T, D, code, thr, 1.0, 1.0, 1.0,
              wpay.data_ptr(), nws, 0, _lib.stream_ptr())
    one = torch.ones(1, device="cuda")
    _lib.call("cfa_sparc_loss_bwd", v.data_ptr(), l.data_ptr(), mask.data_ptr(), B, P, T, D, code, thr, 1.0, 1.0, 1.0,
              wpay.data_ptr(), nws, 0, 0, one.data_ptr(), 0, 0, 0, 0, dpay.data_ptr(), epay.data_ptr(), 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    for name, buf, nbytes in (("v", vb, B * P * D * esz), ("l", lb, B * T * D * esz), ("workspace", wb, nws),
                              ("dv", db, B * P * D * esz), ("dl", eb, B * T * D * esz)):
        n = (nbytes + 127) & ~127
        assert bool((buf[:guard] == 0xA5).all()), f"{name}: bytes written BEFORE the buffer"
        assert bool((buf[guard + n:] == 0xA5).all()), f"{name}: bytes written AFTER the buffer"
        if n > nbytes:
            assert bool((buf[guard + nbytes:guard + n] == 0xA5).all()), f"{name}: bytes written in the alignment tail"
    assert torch.equal(v, v0) and torch.equal(l, l0)
    dv = dpay.view(dtype).view(B, P, D); dl = epay.view(dtype).view(B, T, D)
    assert torch.isfinite(dv.float()).all() and torch.isfinite(dl.float()).all()
    assert float(dv.float().abs().sum()) > 0 and float(dl.float().abs().sum()) > 0


@pytest.mark.parametrize("B", [64, 200, 256, 300])
def test_overlapped_backward_is_race_free(B):
    """The one-call backward launches sparc_bwd3 with programmatic dependent launch under the global InfoNCE backward and
    waits for `d pooled` only before its output pass (csrc/common.cuh).  A wait in the wrong place would read stale
    gradients intermittently, at batch sizes where the two grids really run side by side: 20 repetitions must reproduce,
    bit for bit, the gradients of the per-stage entry points (ordinary stream order, same arithmetic)."""
    from clip_finegrained_alignment_b200 import SPARCLoss
    P, T, D = 196, 77, 512
    g = torch.Generator().manual_seed(B)
    v0 = torch.randn(B, P, D, generator=g).to(torch.bfloat16).cuda()
    l0 = torch.randn(B, T, D, generator=g).to(torch.bfloat16).cuda()
    m = torch.ones(B, T, dtype=torch.bool, device="cuda")
    c = _cfg(1.0 / P, 0.8, 1.2, 2.0)
    def run(fused):
        v = v0.clone().requires_grad_(True); l = l0.clone().requires_grad_(True)
        out = SPARCLoss(c, fused_calls=fused)(v, l, m)
        (out["total_loss"] * 3.0 + out["loss_lv"]).backward()
        return v.grad, l.grad
    rv, rl = run(False)
    torch.cuda.synchronize()
    junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for it in range(20):
        if it % 3 == 0:
            junk.random_(0, 255)                          # different cache / timing state between repetitions
        dv, dl = run(True)
        if B <= 256:
            assert torch.equal(dv, rv) and torch.equal(dl, rl), it
        else:
            # more than two column tiles: the global backward adds its per-tile contributions with red.global (order not
            # fixed) -> last-bit differences in d pooled, with or without the overlap; a stale read would be O(1)
            assert rel_err(dv.float(), rv.float()) <= 1e-4 and rel_err(dl.float(), rl.float()) <= 1e-4, it
