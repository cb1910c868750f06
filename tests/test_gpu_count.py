"""Counting losses on the GPU (CountLoss, CLIPCountLoss) vs reference-made fixtures and the oracle."""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu


def test_count_loss_vs_reference_golden_fp32():
    from clip_finegrained_alignment_b200 import CountLoss
    f = load_golden("countloss_b6_c5_d32.pt")
    ref = f["torch.float64"]
    xs = {k: v.float().cuda().requires_grad_(True) for k, v in f["inputs"].items()}
    out = CountLoss(f["T"], f["alpha"])(xs["la"], xs["lb"], xs["ei"], xs["ek"], xs["cf"])
    out["total_loss"].backward()
    for k in ("clip_loss", "count_loss", "total_loss"):
        assert abs(float(out[k]) - float(ref["losses"][k])) <= 1e-5 * abs(float(ref["losses"][k])), k
    for k in ("la", "lb", "ei", "ek", "cf"):
        assert rel_err(xs[k].grad, ref["grads"][k]) <= 1e-5, (k, rel_err(xs[k].grad, ref["grads"][k]))
    # the reference's own fp32 run agrees to the same level
    r32 = f["torch.float32"]
    assert abs(float(out["total_loss"]) - float(r32["losses"]["total_loss"])) <= 2e-5 * abs(float(r32["losses"]["total_loss"]))


@pytest.mark.parametrize("B,C,D,T,dtype,rtol", [(37, 9, 512, 0.07, torch.float32, 1e-5), (8, 1, 64, 1.0, torch.float32, 1e-5),
                                                (16, 30, 768, 0.07, torch.bfloat16, 2 ** -7)])
def test_count_contrastive_sizes_vs_oracle(B, C, D, T, dtype, rtol):
    from clip_finegrained_alignment_b200.losses import _CountContrastiveFunction
    g = torch.Generator().manual_seed(B + C)
    ei = torch.randn(B, D, generator=g).to(dtype); ek = torch.randn(B, D, generator=g).to(dtype)
    cf = torch.randn(B, C, D, generator=g).to(dtype)
    for include_pos in (False, True):
        fw = lo.count_contrastive_forward(ei.double(), ek.double(), cf.double(), T, include_pos)
        ref = lo.count_contrastive_backward(fw, 2.5)
        xs = [t.cuda().requires_grad_(True) for t in (ei, ek, cf)]
        loss = _CountContrastiveFunction.apply(*xs, T, include_pos)
        (2.5 * loss).backward()
        assert abs(float(loss) - float(fw["loss"])) <= 1e-5 * max(1.0, abs(float(fw["loss"])))
        for x, r in zip(xs, ref):
            assert rel_err(x.grad.float(), r) <= rtol, (include_pos, rel_err(x.grad.float(), r))


def test_clip_count_loss_vs_reference_golden():
    from clip_finegrained_alignment_b200 import CLIPCountLoss
    f = load_golden("clipcount_b4_t3_d24.pt")
    ref = f["torch.float64"]
    img = f["img"].float().cuda().requires_grad_(True)
    txt = f["txt"].float().cuda().requires_grad_(True)
    crit = CLIPCountLoss(0.07, 0.5)
    out = crit(img, txt, torch.arange(12, device="cuda"))
    out["total_loss"].backward()
    assert abs(float(out["clip_loss"]) - float(ref["losses"]["clip_loss"])) <= 1e-5 * float(ref["losses"]["clip_loss"])
    assert float(out["count_loss"]) == 0.0 and out["count_loss"].dtype == torch.float64      # losses.py:55: fp64 zero
    assert out["total_loss"].dtype == f["torch.float32"]["losses"]["total_loss"].dtype
    assert rel_err(img.grad, ref["da"]) <= 2e-5 and rel_err(txt.grad, ref["db"]) <= 2e-5
    out2 = crit(img, txt)                                     # no count features: fp32 zero (losses.py:117)
    assert float(out2["count_loss"]) == 0.0 and out2["total_loss"].dtype == torch.float32
    with pytest.raises(IndexError):                           # the reference's behaviour for any other group size
        crit(img, txt, torch.arange(24, device="cuda"))


def test_count_loss_rejects_cpu_and_bad_shapes():
    from clip_finegrained_alignment_b200 import CountLoss, _lib
    with pytest.raises(_lib.CfaError):
        CountLoss()(torch.randn(3, 3), torch.randn(3, 3), torch.randn(3, 4), torch.randn(3, 4), torch.randn(3, 2, 4))
    x = torch.randn(3, 3, device="cuda")
    with pytest.raises(_lib.CfaError):
        CountLoss()(x, x, torch.randn(3, 4, device="cuda"), torch.randn(3, 4, device="cuda"), torch.randn(2, 2, 4, device="cuda"))
