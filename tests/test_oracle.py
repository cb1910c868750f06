"""The oracle (oracle/*.py) against the golden fixtures made by the reference itself,
and against the live reference when /root/reference is present.  CPU only."""
import glob
import os
import types

import pytest
import torch

from conftest import GOLDEN, load_golden, reference_modules, rel_err
from oracle import adamspd_oracle as ao
from oracle import losses_oracle as lo

SPARC_FIX = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "sparc_b*.pt")))
CLIP_FIX = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "clip_*.pt")))
ADAM_FIX = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN, "adamspd_*.pt")))


def test_fixtures_present():
    assert len(SPARC_FIX) >= 4 and len(CLIP_FIX) >= 2 and len(ADAM_FIX) >= 3


@pytest.mark.parametrize("fname", SPARC_FIX)
def test_sparc_oracle_matches_reference_fp64(fname):
    f = load_golden(fname)
    v, l = f["v"].double(), f["l"].double()
    out = lo.sparc_forward(v, l, f["mask"], f["thr"], f["gw"], f["lw"], f["s"])
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(f["f64"]["losses"][k])) <= 1e-12 * max(1.0, abs(float(out[k]))), k
    dv, dl = lo.sparc_backward(out, {f["backprop"]: 1.0})
    assert rel_err(dv, f["f64"]["dv"]) < 1e-10
    assert rel_err(dl, f["f64"]["dl"]) < 1e-10


@pytest.mark.parametrize("fname", [n for n in SPARC_FIX if "thr05" not in n])
def test_sparc_oracle_matches_reference_fp32(fname):
    # thr = 1/P: well conditioned, fp32 reference and fp32 oracle agree to rounding (SURVEY finding 2)
    f = load_golden(fname)
    out = lo.sparc_forward(f["v"], f["l"], f["mask"], f["thr"], f["gw"], f["lw"], f["s"])
    for k in lo.SPARC_KEYS:
        torch.testing.assert_close(out[k], f["f32"]["losses"][k], rtol=1e-5, atol=1e-6)
    dv, dl = lo.sparc_backward(out, {f["backprop"]: 1.0})
    assert rel_err(dv, f["f32"]["dv"]) < 1e-5
    assert rel_err(dl, f["f32"]["dl"]) < 1e-5


def test_sparc_truncate_equals_reference_for_full_mask():
    f = load_golden(SPARC_FIX[0])
    v, l = f["v"].double(), f["l"].double()
    a = lo.sparc_forward(v, l, f["mask"], f["thr"], 1.0, 1.0, f["s"], mask_semantics="reference")
    b = lo.sparc_forward(v, l, f["mask"], f["thr"], 1.0, 1.0, f["s"], mask_semantics="truncate")
    for k in lo.SPARC_KEYS:
        assert float(a[k]) == float(b[k])


def test_sparc_masked_truncate_semantics_pinned():
    """Padded masks: the reference's local loss is NaN (SURVEY finding 3); the 'truncate'
    semantics equal the reference evaluated per sample on valid tokens only."""
    f = load_golden("sparc_masked_b4_p50_d64.pt")
    assert f["ref_local_is_nan"]
    v, l = f["v"].double(), f["l"].double()
    ref = lo.sparc_forward(v, l, f["mask"], f["thr"], 1.0, 1.0, f["s"], mask_semantics="reference")
    assert torch.isnan(ref["local_loss"])
    out = lo.sparc_forward(v, l, f["mask"], f["thr"], 1.0, 1.0, f["s"], mask_semantics="truncate")
    assert abs(float(out["loss_vl_local"]) - float(f["loss_vl_local"])) < 1e-6
    assert abs(float(out["loss_lv_local"]) - float(f["loss_lv_local"])) < 1e-6
    assert abs(float(out["global_loss"]) - float(f["global_loss"])) < 1e-6
    dv, dl = lo.sparc_backward(out, {"total_loss": 1.0})
    assert rel_err(dv, f["dv"]) < 1e-5      # fixture inputs were stored as fp32
    assert rel_err(dl, f["dl"]) < 1e-5


def test_sparc_weights_sum_to_one_and_symmetry():
    g = torch.Generator().manual_seed(3)
    v = torch.randn(3, 40, 16, generator=g, dtype=torch.float64)
    l = torch.randn(3, 9, 16, generator=g, dtype=torch.float64)
    m = torch.ones(3, 9, dtype=torch.bool)
    out = lo.sparc_forward(v, l, m, 1 / 40, 1.0, 1.0, 2.0)
    W = out["_cache"]["W"]
    torch.testing.assert_close(W.sum(-1), torch.ones(3, 9, dtype=torch.float64))
    # both global directions come from one logits matrix (losses.py:215-216)
    torch.testing.assert_close(out["_cache"]["g1"]["logits"], out["_cache"]["g2"]["logits"].t())


@pytest.mark.parametrize("key", lo.SPARC_KEYS)
def test_sparc_backward_every_output_vs_autograd(key):
    """Each of the 7 dict entries is differentiable; manual backward == autograd of the oracle forward."""
    g = torch.Generator().manual_seed(11)
    v = torch.randn(2, 21, 12, generator=g, dtype=torch.float64, requires_grad=True)
    l = torch.randn(2, 7, 12, generator=g, dtype=torch.float64, requires_grad=True)
    m = torch.ones(2, 7, dtype=torch.bool)
    out = lo.sparc_forward(v, l, m, 1 / 21, 0.8, 1.2, 3.0)
    out[key].backward()
    with torch.no_grad():
        out2 = lo.sparc_forward(v.detach(), l.detach(), m, 1 / 21, 0.8, 1.2, 3.0)
        dv, dl = lo.sparc_backward(out2, {key: 1.0})
    assert rel_err(dv, v.grad) < 1e-10
    assert rel_err(dl, l.grad) < 1e-10


@pytest.mark.parametrize("fname", CLIP_FIX)
def test_clip_oracle(fname):
    f = load_golden(fname)
    a, b = f["a"].double(), f["b"].double()
    out = lo.clip_loss_forward(a, b, f["temperature"])
    assert abs(float(out["clip_loss"]) - float(f["f64"]["clip_loss"])) < 1e-12
    da, db = lo.clip_loss_backward(out, f["temperature"])
    assert rel_err(da, f["f64"]["da"]) < 1e-10
    assert rel_err(db, f["f64"]["db"]) < 1e-10
    out32 = lo.clip_loss_forward(f["a"], f["b"], f["temperature"])
    torch.testing.assert_close(out32["clip_loss"], f["f32"]["clip_loss"], rtol=1e-5, atol=1e-6)


def test_pairwise_oracle():
    f = load_golden("pairwise_b12_d32.pt")
    a, b = f["a"].double(), f["b"].double()
    fw = lo.infonce_forward(a, b, f["s"])
    assert abs(float(fw["loss_sum"]) / 12 - float(f["loss"])) < 1e-6   # fixture inputs stored fp32
    z = torch.zeros(12, dtype=torch.float64)
    da, db = lo.symmetric_infonce_backward(fw["ah"], fw["an"], fw["bh"], fw["bn"], fw["lse"], z, f["s"],
                                           1.0, 0.0, 12.0)
    assert rel_err(da, f["da"]) < 1e-5
    assert rel_err(db, f["db"]) < 1e-5


def test_masked_pairwise_oracle():
    """masked_pairwise_contrastive_loss (losses.py:165-197) vs the reference: all-True mask, and the padded mask under
    'truncate' semantics = the reference evaluated per sample on its valid tokens (the reference itself is NaN there)."""
    f = load_golden("maskedpair_b3_t20_d48.pt")
    a, b = f["a"].double(), f["b"].double()
    full = torch.ones(a.shape[:2], dtype=torch.bool)
    for sem in ("reference", "truncate"):
        fw = lo.masked_pairwise_forward(a, b, full, f["s"], mask_semantics=sem)
        assert abs(float(fw["loss"]) - float(f["loss"])) < 1e-6
        da, db = lo.masked_pairwise_backward(fw)
        assert rel_err(da, f["da"]) < 1e-5 and rel_err(db, f["db"]) < 1e-5
    assert f["ref_padded_is_nan"]
    assert torch.isnan(lo.masked_pairwise_forward(a, b, f["mask"], f["s"], mask_semantics="reference")["loss"])
    fw = lo.masked_pairwise_forward(a, b, f["mask"], f["s"])
    assert abs(float(fw["loss"]) - float(f["loss_trunc"])) < 1e-6
    da, db = lo.masked_pairwise_backward(fw)
    assert rel_err(da, f["da_trunc"]) < 1e-5 and rel_err(db, f["db_trunc"]) < 1e-5


def test_gathered_infonce_equals_concatenated():
    """SURVEY §8e oracle: N ranks with gathered columns == single process on the concatenated batch."""
    g = torch.Generator().manual_seed(5)
    N, B, D, s = 4, 6, 16, 3.0
    a = torch.randn(N * B, D, generator=g, dtype=torch.float64)
    b = torch.randn(N * B, D, generator=g, dtype=torch.float64)
    f1 = lo.infonce_forward(a, b, s)
    f2 = lo.infonce_forward(b, a, s)
    ref_loss = 0.5 * (f1["loss_sum"] + f2["loss_sum"]) / (N * B)
    da_ref, db_ref = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"],
                                                   s, 0.5, 0.5, float(N * B))
    per_rank = [lo.gathered_infonce_rank(a[r * B:(r + 1) * B], b[r * B:(r + 1) * B], a, b, r, s, 0.5, 0.5)
                for r in range(N)]
    loss = sum(0.5 * (fa["loss_sum"] + fb["loss_sum"]) for fa, fb, _ in per_rank) / (N * B)
    assert abs(float(loss - ref_loss)) < 1e-12
    lse_a = torch.cat([fa["lse"] for fa, _, _ in per_rank])
    lse_b = torch.cat([fb["lse"] for _, fb, _ in per_rank])
    for r, (_, _, bwd) in enumerate(per_rank):
        da, db = bwd(lse_a, lse_b)
        assert rel_err(da, da_ref[r * B:(r + 1) * B]) < 1e-11
        assert rel_err(db, db_ref[r * B:(r + 1) * B]) < 1e-11


@pytest.mark.parametrize("fname", ADAM_FIX)
def test_adamspd_oracle_matches_reference(fname):
    f = load_golden(fname)
    p = [x.clone() for x in f["p0"]]
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    vmax = [torch.zeros_like(x) for x in p] if f["amsgrad"] else None
    steps = [0] * len(p)
    took = 0
    for t in range(f["steps"]):
        grads = [None if (j in f["none_grad_idx"] and t % 2 == 1) else f["grads"][t][j] for j in range(len(p))]
        st = ao.adamspd_step(p, grads, m, v, f["pre"], steps, f["lr"], f["betas"], f["eps"], f["wd"], vmax)
        took += sum(1 for s in st if s is not None and s[0])
        if t + 1 in f["snaps"]:
            for x, r in zip(p, f["snaps"][t + 1]):
                assert torch.equal(x, r), (fname, t)     # same ops, same order -> bit-exact on CPU
    for j, s in enumerate(f["state"]):
        assert s["step"] == steps[j]
        assert torch.equal(s["exp_avg"], m[j]) and torch.equal(s["exp_avg_sq"], v[j])
    if f["pre"] is not None:
        assert took > 0          # the SPD branch (optimizers.py:148-150) was exercised


def test_count_losses_oracle_matches_reference():
    """CountLoss / CLIPCountLoss restatement vs reference outputs and autograd gradients (fp64: 1e-11)."""
    f = load_golden("countloss_b6_c5_d32.pt")
    x = f["inputs"]
    ref = f["torch.float64"]
    fw = lo.count_loss_forward(x["la"], x["lb"], x["ei"], x["ek"], x["cf"], f["T"], f["alpha"])
    for k in ("clip_loss", "count_loss", "total_loss"):
        assert abs(float(fw[k]) - float(ref["losses"][k])) < 1e-12
    dla, dlb = lo.logits_ce_backward(x["la"], x["lb"], fw["_f1"])
    dei, dek, dcf = lo.count_contrastive_backward(fw["_f2"], f["alpha"])
    for got, key in ((dla, "la"), (dlb, "lb"), (dei, "ei"), (dek, "ek"), (dcf, "cf")):
        assert rel_err(got, ref["grads"][key]) < 1e-11, key
    f = load_golden("clipcount_b4_t3_d24.pt")
    ref = f["torch.float64"]
    fw = lo.clip_count_forward(f["img"], f["txt"], f["T"])
    assert abs(float(fw["clip_loss"]) - float(ref["losses"]["clip_loss"])) < 1e-12
    assert float(ref["losses"]["count_loss"]) == 0.0 and float(fw["count_loss"]) == 0.0
    da, db = lo.clip_count_backward(fw, f["T"])
    assert rel_err(da, ref["da"]) < 1e-11 and rel_err(db, ref["db"]) < 1e-11
    # the positive-in-denominator variant is what CLIPCountLoss.count_loss computes per group (losses.py:78-86)
    g = torch.Generator().manual_seed(3)
    ei = torch.randn(3, 8, generator=g, dtype=torch.float64); ek = torch.randn(3, 8, generator=g, dtype=torch.float64)
    cf = torch.randn(3, 4, 8, generator=g, dtype=torch.float64)
    fw = lo.count_contrastive_forward(ei, ek, cf, 0.07, include_pos=True)
    n = lambda t: t / t.norm(dim=-1, keepdim=True)
    want = 0.0
    for i in range(3):
        pos = torch.dot(n(ei)[i], n(ek)[i]); neg = n(cf)[i] @ n(ei)[i]
        num = torch.exp(pos / 0.07); want += -torch.log(num / (num + torch.exp(neg / 0.07).sum()))
    assert abs(float(fw["loss"]) - float(want / 3)) < 1e-12


def test_amp_step_oracle_matches_reference_sequence():
    """unscale_ -> clip_grad_norm_ -> scaler.step -> update (finetuner.py:150-153), restated, vs the fixture made with
    the reference AdamSPD and torch's GradScaler: bit-exact parameters, norms, scales, skipped steps."""
    f = load_golden("ampstep_s16.pt")
    p = [x.clone() for x in f["p0"]]
    m = [torch.zeros_like(x) for x in p]
    v = [torch.zeros_like(x) for x in p]
    steps = [0] * len(p)
    scale, tracker = 1024.0, 0
    for t in range(f["steps"]):
        assert scale == f["scales"][t]
        grads = [g * scale for g in f["grads"][t]]
        if t in f["inf_steps"]:
            grads[2].view(-1)[3] = float("inf")
        eff, total, found = ao.amp_unscale_clip(grads, scale, f["max_norm"])
        assert found == f["skipped"][t]
        if not found:
            assert abs(total - f["norms"][t]) <= 1e-6 * f["norms"][t]
            ao.adamspd_step(p, eff, m, v, f["pre"], steps, f["lr"], (0.9, 0.999), 1e-8, f["wd"])
        scale, tracker = ao.amp_scale_update(scale, tracker, found, growth_interval=3)
        if t + 1 in f["snaps"]:
            for x, r in zip(p, f["snaps"][t + 1]):
                assert torch.equal(x, r), t
    assert steps == f["state_steps"] and scale == f["final_scale"]


def test_live_reference_if_present():
    mods = reference_modules()
    if mods is None:
        pytest.skip("reference not mounted (GPU box)")
    ref_losses, _ = mods
    g = torch.Generator().manual_seed(21)
    v = torch.randn(3, 50, 32, generator=g, dtype=torch.float64, requires_grad=True)
    l = torch.randn(3, 77, 32, generator=g, dtype=torch.float64, requires_grad=True)
    m = torch.ones(3, 77, dtype=torch.bool)
    cfg = types.SimpleNamespace(similarity_threshold=1 / 50, global_loss_weight=1.0, local_loss_weight=1.0,
                                inverse_temperature=1.0)
    ref = ref_losses.SPARCLoss(cfg)(v, l, m)
    ref["total_loss"].backward()
    out = lo.sparc_forward(v.detach(), l.detach(), m, 1 / 50, 1.0, 1.0, 1.0)
    for k in lo.SPARC_KEYS:
        assert abs(float(out[k]) - float(ref[k])) < 1e-12
    dv, dl = lo.sparc_backward(out)
    assert rel_err(dv, v.grad) < 1e-10 and rel_err(dl, l.grad) < 1e-10
