"""AdamSPD CUDA kernels (through the C ABI) vs the oracle and the reference-made golden fixtures."""
import pytest
import torch

from conftest import load_golden
from oracle import adamspd_oracle as ao

pytestmark = pytest.mark.gpu


def _run_case(f, device="cuda"):
    from clip_finegrained_alignment_b200 import AdamSPD
    params = [torch.nn.Parameter(x.clone().to(device)) for x in f["p0"]]
    pre = None if f["pre"] is None else [x.clone().to(device) for x in f["pre"]]
    opt = AdamSPD([{"params": params, "pre": pre}], lr=f["lr"], betas=f["betas"], eps=f["eps"], weight_decay=f["wd"],
                  amsgrad=f["amsgrad"])
    snaps = {}
    for t in range(f["steps"]):
        for j, p in enumerate(params):
            p.grad = None if (j in f["none_grad_idx"] and t % 2 == 1) else f["grads"][t][j].clone().to(device)
        opt.step()
        if t + 1 in f["snaps"]:
            snaps[t + 1] = [p.detach().cpu().clone() for p in params]
    return opt, params, snaps


@pytest.mark.parametrize("name", ["adamspd_s20.pt", "adamspd_s12_ams_lr1e3.pt", "adamspd_s8_nopre_nonegrad.pt"])
def test_adamspd_matches_reference_golden(name):
    f = load_golden(name)
    opt, params, snaps = _run_case(f)
    for step, ref in f["snaps"].items():
        for x, r in zip(snaps[step], ref):
            # north-star bar: parameters within 1e-6 (absolute) of optimizers.py
            assert (x - r).abs().max().item() <= 1e-6, (name, step)
    for p, s in zip(params, f["state"]):
        st = opt.state[p]
        assert st["step"] == s["step"]
        assert set(st.keys()) >= {"step", "exp_avg", "exp_avg_sq", "hyper"}
        torch.testing.assert_close(st["exp_avg"].cpu(), s["exp_avg"], rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(st["exp_avg_sq"].cpu(), s["exp_avg_sq"], rtol=1e-5, atol=1e-12)
        if f["amsgrad"]:
            torch.testing.assert_close(st["max_exp_avg_sq"].cpu(), s["max_exp_avg_sq"], rtol=1e-5, atol=1e-12)


def test_adamspd_100_steps_vs_oracle_both_branches():
    """100 steps, random grads so both SPD branches run, multi-chunk and ragged tensors,
    reference hyper-parameters (finetuner.py:297-318).  Bar: |p - p_ref| <= 1e-6."""
    from clip_finegrained_alignment_b200 import AdamSPD
    g = torch.Generator().manual_seed(1234)
    sizes = [(1,), (5,), (8192,), (8193,), (3, 8192), (257, 129), (70001,)]
    p0 = [torch.randn(*s, generator=g) * 0.02 for s in sizes]
    pre = [p + 1e-3 * torch.randn(*p.shape, generator=g) for p in p0]
    lr, betas, eps, wd = 2e-5, (0.9, 0.999), 1e-8, 0.1
    ref_p = [x.clone() for x in p0]
    m = [torch.zeros_like(x) for x in p0]
    v = [torch.zeros_like(x) for x in p0]
    steps = [0] * len(p0)
    params = [torch.nn.Parameter(x.clone().cuda()) for x in p0]
    opt = AdamSPD([{"params": params, "pre": [x.cuda() for x in pre]}], lr=lr, betas=betas, eps=eps, weight_decay=wd)
    n_proj = n_ratio = agree = total = 0
    for t in range(100):
        grads = [torch.randn(*s, generator=g) * 1e-3 for s in sizes]
        for p, gr in zip(params, grads):
            p.grad = gr.cuda()
        opt.step()
        st = ao.adamspd_step(ref_p, grads, m, v, pre, steps, lr, betas, eps, wd)
        dev_stats = opt.last_step_stats.cpu()
        for j, (proj, ratio) in enumerate(st):
            total += 1
            agree += int(bool(dev_stats[j, 0].item()) == proj)
            n_proj += int(proj)
            n_ratio += int(ratio > 0)
    assert n_proj > 50 and n_ratio > 20 and n_proj < total     # both branches exercised
    assert agree == total                                     # same per-tensor decisions as the reference
    for x, r in zip(params, ref_p):
        assert (x.detach().cpu() - r).abs().max().item() <= 1e-6


def test_adamspd_wd0_equals_adam():
    """Property: with weight_decay = 0 the projection is a no-op, AdamSPD == Adam without decay."""
    from clip_finegrained_alignment_b200 import AdamSPD
    torch.manual_seed(0)
    w = torch.randn(1000, 37, device="cuda")
    p1 = torch.nn.Parameter(w.clone())
    p2 = torch.nn.Parameter(w.clone())
    o1 = AdamSPD([{"params": [p1], "pre": [w.clone() + 0.01]}], lr=1e-3, weight_decay=0.0)
    o2 = torch.optim.Adam([p2], lr=1e-3)
    for _ in range(10):
        g = torch.randn_like(w)
        p1.grad = g.clone()
        p2.grad = g.clone()
        o1.step()
        o2.step()
    torch.testing.assert_close(p1, p2, rtol=1e-5, atol=1e-6)


def test_adamspd_state_dict_roundtrip_and_unaligned_views():
    from clip_finegrained_alignment_b200 import AdamSPD
    torch.manual_seed(1)
    base = torch.randn(10001, device="cuda")
    p = torch.nn.Parameter(base[1:])                 # 4-byte aligned only: scalar path
    pre = [base[1:].clone()]
    opt = AdamSPD([{"params": [p], "pre": pre}], lr=1e-3, weight_decay=0.1)
    ref = p.detach().cpu().clone()
    m, v, steps = [torch.zeros_like(ref)], [torch.zeros_like(ref)], [0]
    for _ in range(3):
        g = torch.randn(10000, device="cuda") * 1e-2
        p.grad = g
        opt.step()
        ao.adamspd_step([ref], [g.cpu()], m, v, [pre[0].cpu()], steps, 1e-3, (0.9, 0.999), 1e-8, 0.1)
    assert (p.detach().cpu() - ref).abs().max().item() <= 1e-6
    sd = opt.state_dict()
    assert "pre" in sd["param_groups"][0] and sd["state"][0]["step"] == 3
    opt2 = AdamSPD([{"params": [p], "pre": pre}], lr=1e-3, weight_decay=0.1)
    opt2.load_state_dict(sd)
    assert opt2.state[p]["step"] == 3
    p.grad = torch.randn(10000, device="cuda") * 1e-2
    opt2.step()
    assert opt2.state[p]["step"] == 4


def test_amp_step_matches_reference_sequence_golden():
    """AdamSPD.amp_step(scaler, max_norm) == scaler.unscale_ + clip_grad_norm_ + scaler.step of the reference
    (finetuner.py:150-152; fixture: reference AdamSPD + torch GradScaler on CPU).  Covers clipped and unclipped steps,
    two skipped (inf) steps with scale back-off, scale growth, and the host step counters.  Bar: |p - p_ref| <= 1e-6."""
    from clip_finegrained_alignment_b200 import AdamSPD
    f = load_golden("ampstep_s16.pt")
    params = [torch.nn.Parameter(x.clone().cuda()) for x in f["p0"]]
    opt = AdamSPD([{"params": params, "pre": [x.cuda() for x in f["pre"]]}], lr=f["lr"], betas=(0.9, 0.999), eps=1e-8,
                  weight_decay=f["wd"])
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0, growth_interval=3)
    scaler.scale(torch.zeros(1, device="cuda"))
    for t in range(f["steps"]):
        scale = float(scaler.get_scale())
        assert scale == f["scales"][t], (t, scale)
        for j, p in enumerate(params):
            p.grad = (f["grads"][t][j] * scale).cuda()
        if t in f["inf_steps"]:
            params[2].grad.view(-1)[3] = float("inf")
        raw = [p.grad.clone() for p in params]
        total = opt.amp_step(scaler, f["max_norm"])
        scaler.update()
        for p, r in zip(params, raw):                      # gradients are left untouched
            assert torch.equal(p.grad, r) or t in f["inf_steps"]
        if not f["skipped"][t]:
            assert abs(float(total) - f["norms"][t]) <= 1e-5 * f["norms"][t], (t, float(total), f["norms"][t])
        if t + 1 in f["snaps"]:
            for x, r in zip(params, f["snaps"][t + 1]):
                assert (x.detach().cpu() - r).abs().max().item() <= 1e-6, t
    sd = opt.state_dict()                                   # settles the device-side skip decisions
    assert [sd["state"][j]["step"] for j in range(len(params))] == f["state_steps"]
    assert float(scaler.get_scale()) == f["final_scale"]


def test_amp_step_without_scaler_equals_clip_then_step():
    """No GradScaler (fp32 training, finetuner.py:156-183 path): amp_step(None, max_norm) == clip_grad_norm_ + step."""
    from clip_finegrained_alignment_b200 import AdamSPD
    torch.manual_seed(5)
    shapes = [(70001,), (300, 41), (8192,)]
    w = [torch.randn(*s, device="cuda") * 0.02 for s in shapes]
    pre = [x + 1e-3 * torch.randn_like(x) for x in w]
    pa = [torch.nn.Parameter(x.clone()) for x in w]
    pb = [torch.nn.Parameter(x.clone()) for x in w]
    oa = AdamSPD([{"params": pa, "pre": pre}], lr=2e-5, weight_decay=0.1)
    ob = AdamSPD([{"params": pb, "pre": pre}], lr=2e-5, weight_decay=0.1)
    for t in range(6):
        gs = [torch.randn(*s, device="cuda") * (1e-3 if t % 2 else 1e-1) for s in shapes]
        for p, q, g in zip(pa, pb, gs):
            p.grad = g.clone(); q.grad = g.clone()
        na = oa.amp_step(None, 1.0)
        nb = torch.nn.utils.clip_grad_norm_(pb, 1.0)
        ob.step()
        assert abs(float(na) - float(nb)) <= 1e-5 * float(nb)
    for p, q in zip(pa, pb):
        assert (p - q).abs().max().item() <= 1e-7
