"""Oracle comparisons at the BASELINE configurations' own sizes (parity gap (i)):
  config 2  B = 256, P = 196, T = 77, D = 512, bf16   SPARCLoss through the public class vs the fp64 oracle
  config 3  one rank's all-gathered global InfoNCE, local 1024 rows x 8192 gathered columns, D = 512 (fp32 oracle)
  config 5  ONE AdamSPD step over the 590-tensor ViT-L/14 CLIP list vs the oracle
and the north-star gradient bar for bf16 outputs: the kernel's bf16 gradient against the oracle gradient ROUNDED to
bf16, element-wise within one bf16 ulp for >= 99.9 % of the elements (an error of 4e-3 in the pre-rounding value would
move far more than 0.1 % of the elements by more than one ulp; the relative-Frobenius bar alone cannot tell)."""
import types

import pytest
import torch

from conftest import rel_err
from oracle import adamspd_oracle as ao
from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu


def _cfg(thr, gw=1.0, lw=1.0, s=1.0):
    return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=gw, local_loss_weight=lw,
                                 inverse_temperature=s)


def bf16_ulp_outliers(got_bf16: torch.Tensor, ref_f64: torch.Tensor) -> float:
    """Fraction of elements whose bf16 value differs from round_bf16(ref) by more than one bf16 ulp of round_bf16(ref)."""
    ref_b = ref_f64.float().to(torch.bfloat16)
    r = ref_b.float().cpu()
    g = got_bf16.float().cpu()
    # one ulp of a bf16 number x: 2^(floor(log2|x|) - 7); subnormal range is irrelevant for gradients of this size
    ulp = torch.pow(2.0, torch.floor(torch.log2(r.abs().clamp_min(1e-30))) - 7)
    return float(((g - r).abs() > ulp * 1.0001).double().mean())


def test_config2_full_size_vs_oracle():
    from clip_finegrained_alignment_b200 import SPARCLoss
    B, P, T, D = 256, 196, 77, 512
    g = torch.Generator().manual_seed(42)
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool)
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    vv = v.cuda().requires_grad_(True)
    ll = l.cuda().requires_grad_(True)
    out = SPARCLoss(_cfg(thr))(vv, ll, m.cuda())
    out["total_loss"].backward()
    torch.cuda.synchronize()
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 1.0, 1.0, 1.0)
    rv, rl = lo.sparc_backward(o)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(o[k])) <= 1e-4 * max(1.0, abs(float(o[k]))), (k, float(out[k]), float(o[k]))
    assert rel_err(vv.grad.float(), rv) <= 1e-3 + 2.0 ** -8
    assert rel_err(ll.grad.float(), rl) <= 1e-3 + 2.0 ** -8
    # north-star bar on bf16 outputs: <= 1 ulp of the bf16-rounded oracle gradient for >= 99.9 % of the elements
    assert bf16_ulp_outliers(vv.grad, rv) <= 1e-3, bf16_ulp_outliers(vv.grad, rv)
    assert bf16_ulp_outliers(ll.grad, rl) <= 1e-3, bf16_ulp_outliers(ll.grad, rl)


@pytest.mark.parametrize("B,P,T,D,s", [(3, 196, 77, 512, 1.0), (2, 197, 77, 512, 2.0), (4, 50, 77, 256, 1.0)])
def test_tc_gradients_within_one_bf16_ulp(B, P, T, D, s):
    from clip_finegrained_alignment_b200 import SPARCLoss
    g = torch.Generator().manual_seed(B * 7 + P)
    v = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    l = torch.randn(B, T, D, generator=g).to(torch.bfloat16)
    m = torch.ones(B, T, dtype=torch.bool)
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    vv = v.cuda().requires_grad_(True)
    ll = l.cuda().requires_grad_(True)
    SPARCLoss(_cfg(thr, 0.9, 1.1, s), kernel_path="tc")(vv, ll, m.cuda())["total_loss"].backward()
    o = lo.sparc_forward(v.double(), l.double(), m, thr, 0.9, 1.1, s)
    rv, rl = lo.sparc_backward(o)
    assert bf16_ulp_outliers(vv.grad, rv) <= 1e-3, bf16_ulp_outliers(vv.grad, rv)
    assert bf16_ulp_outliers(ll.grad, rl) <= 1e-3, bf16_ulp_outliers(ll.grad, rl)


def test_config3_gathered_infonce_one_rank_of_eight():
    """Rank 3 of 8: local [1024, 512] rows against the gathered [8192, 512] columns (the shape every GPU of BASELINE
    config 3 works on), tensor-core logits kernels through the C ABI vs the fp32 CPU oracle."""
    from clip_finegrained_alignment_b200 import _lib
    N, B, D, s, rank = 8, 1024, 512, 1.0, 3
    Bg = N * B
    g = torch.Generator().manual_seed(8192)
    a = torch.randn(Bg, D, generator=g)
    b = torch.randn(Bg, D, generator=g)
    ac, bc = a.cuda(), b.cuda()
    assert _lib.lib.cfa_global_infonce_path(B, Bg, D, 0) == 2
    ws_bytes = _lib.lib.cfa_global_infonce_workspace_bytes(B, Bg, D)
    lse, sums, wss, norms = [], [], [], []
    for r in range(N):                                   # every rank's forward: the backward needs all ranks' lse
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        l2 = torch.empty(2, B, device="cuda"); n2 = torch.empty(2, B, device="cuda"); s2 = torch.empty(2, device="cuda")
        al, bl = ac[r * B:(r + 1) * B].contiguous(), bc[r * B:(r + 1) * B].contiguous()
        _lib.call("cfa_global_infonce_fwd", al.data_ptr(), bl.data_ptr(), ac.data_ptr(), bc.data_ptr(), B, Bg, D, r * B, s,
                  1e-12, l2.data_ptr(), n2.data_ptr(), s2.data_ptr(), 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), ws_bytes, 0,
                  0, _lib.stream_ptr())
        lse.append(l2); sums.append(s2); norms.append(n2); wss.append(ws if r == rank else None)
    torch.cuda.synchronize()
    lse_all = torch.cat(lse, dim=1).contiguous()
    # fp32 CPU oracle: one rank's [B, Bg] problems per direction, plus every rank's lse for the cross terms
    per = [lo.gathered_infonce_rank(a[r * B:(r + 1) * B], b[r * B:(r + 1) * B], a, b, r, s, 0.5, 0.5) for r in range(N)]
    lse_a_all = torch.cat([p[0]["lse"] for p in per]); lse_b_all = torch.cat([p[1]["lse"] for p in per])
    torch.testing.assert_close(lse_all[0].cpu(), lse_a_all, rtol=2e-5, atol=2e-5)
    torch.testing.assert_close(lse_all[1].cpu(), lse_b_all, rtol=2e-5, atol=2e-5)
    fa, fb, bwd = per[rank]
    assert abs(float(sums[rank][0]) - float(fa["loss_sum"])) <= 2e-5 * abs(float(fa["loss_sum"]))
    assert abs(float(sums[rank][1]) - float(fb["loss_sum"])) <= 2e-5 * abs(float(fb["loss_sum"]))
    da_ref, db_ref = bwd(lse_a_all, lse_b_all)
    coef = torch.full((2,), 0.5 / Bg, device="cuda")
    sl = slice(rank * B, (rank + 1) * B)
    al, bl = ac[sl].contiguous(), bc[sl].contiguous()
    da = torch.empty(B, D, device="cuda"); db = torch.empty(B, D, device="cuda")
    _lib.call("cfa_global_infonce_bwd", al.data_ptr(), bl.data_ptr(), ac.data_ptr(), bc.data_ptr(), B, Bg, D, rank * B, s,
              1e-12, lse[rank].data_ptr(), lse_all.data_ptr(), norms[rank].data_ptr(), coef.data_ptr(), da.data_ptr(),
              db.data_ptr(), wss[rank].data_ptr(), ws_bytes, 0, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
    assert rel_err(da, da_ref) <= 2e-4, rel_err(da, da_ref)
    assert rel_err(db, db_ref) <= 2e-4, rel_err(db, db_ref)


def test_config5_one_adamspd_step_on_the_vit_l14_list():
    """590 tensors / 427 616 513 fp32 elements (BASELINE config 5): one step, random gradients so that both SPD branches
    occur, parameters within 1e-6 of the oracle (optimizers.py:100-152)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import vit_l14_clip_shapes
    from clip_finegrained_alignment_b200 import AdamSPD
    shapes = vit_l14_clip_shapes()
    g = torch.Generator().manual_seed(5)
    p0 = [torch.randn(*s, generator=g) * 0.02 for s in shapes]
    pre = [x + 1e-3 * torch.randn(*x.shape, generator=g) for x in p0]
    grads = [torch.randn(*x.shape, generator=g) * 1e-3 for x in p0]
    params = [torch.nn.Parameter(x.clone().cuda()) for x in p0]
    for q, gr in zip(params, grads):
        q.grad = gr.cuda()
    opt = AdamSPD([{"params": params, "pre": [x.cuda() for x in pre]}], lr=2e-5, betas=(0.9, 0.999), eps=1e-8,
                  weight_decay=0.1)
    opt.step()
    torch.cuda.synchronize()
    ref = [x.clone() for x in p0]
    stats = ao.adamspd_step(ref, grads, [torch.zeros_like(x) for x in p0], [torch.zeros_like(x) for x in p0], pre,
                            [0] * len(p0), 2e-5, (0.9, 0.999), 1e-8, 0.1)
    worst = max(float((q.detach().cpu() - r).abs().max()) for q, r in zip(params, ref))
    assert worst <= 1e-6, worst
    assert len(stats) == 590
