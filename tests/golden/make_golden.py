"""Generate golden fixtures by running the UNMODIFIED reference
(/root/reference/finetune/{losses,optimizers,config}.py) on seeded inputs.

Run in the build container (the reference is not on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/*.pt (inputs + reference outputs, fp32 stored; fp64 outputs
kept as well for ill-conditioned cases).  The reference is imported, never copied.
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/finetune"


def import_reference():
    sys.path.insert(0, REF)
    import losses as ref_losses          # noqa: E402
    import optimizers as ref_opt         # noqa: E402
    sys.path.pop(0)
    return ref_losses, ref_opt


def cfg(thr, gw, lw, s):
    return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=gw, local_loss_weight=lw,
                                 inverse_temperature=s)


def sparc_case(ref_losses, name, B, P, T, D, thr, gw, lw, s, seed, backprop="total_loss", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, P, D, generator=g, dtype=torch.float32)
    l = torch.randn(B, T, D, generator=g, dtype=torch.float32)
    mask = torch.ones(B, T, dtype=torch.bool)
    outs = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        vv = v.detach().clone().to(dt).requires_grad_(True)
        ll = l.detach().clone().to(dt).requires_grad_(True)
        mod = ref_losses.SPARCLoss(cfg(thr, gw, lw, s))
        out = mod(vv, ll, mask)
        out[backprop].backward()
        outs[tag] = {"losses": {k: o.detach().clone() for k, o in out.items()},
                     "dv": vv.grad.clone(), "dl": ll.grad.clone()}
    fix = dict(name=name, v=v, l=l, mask=mask, thr=thr, gw=gw, lw=lw, s=s, backprop=backprop, **outs)
    torch.save(fix, os.path.join(HERE, f"sparc_{name}.pt"))
    print("sparc", name, {k: float(x) for k, x in outs["f32"]["losses"].items()})


def sparc_masked_case(ref_losses, name, B, P, T, D, thr, s, seed):
    """Pins the 'truncate' mask semantics: reference run per sample on valid tokens only."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from oracle.losses_oracle import sparc_reference_truncated
    g = torch.Generator().manual_seed(seed)
    v = torch.randn(B, P, D, generator=g, dtype=torch.float64)
    l = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    lens = torch.randint(3, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None, :] < lens[:, None]
    mod = ref_losses.SPARCLoss(cfg(thr, 1.0, 1.0, s))
    vv = v.clone().requires_grad_(True)
    ll = l.clone().requires_grad_(True)
    vl, lv = sparc_reference_truncated(mod, vv, ll, mask)
    # global part: the reference is finite for padded masks (losses.py:207-217)
    full = mod(vv, ll, mask)
    total = full["global_loss"] + 0.5 * (vl + lv)
    total.backward()
    dl = ll.grad.clone()
    fix = dict(name=name, v=v.float(), l=l.float(), mask=mask, thr=thr, s=s,
               loss_vl_local=vl.detach(), loss_lv_local=lv.detach(), global_loss=full["global_loss"].detach(),
               loss_vl=full["loss_vl"].detach(), loss_lv=full["loss_lv"].detach(),
               total=total.detach(), dv=vv.grad.clone(), dl=dl,
               ref_local_is_nan=bool(torch.isnan(full["local_loss"])))
    torch.save(fix, os.path.join(HERE, f"sparc_masked_{name}.pt"))
    print("sparc masked", name, float(vl), float(lv), "ref local NaN:", fix["ref_local_is_nan"])


def clip_case(ref_losses, name, B, D, temperature, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, D, generator=g)
    b = torch.randn(B, D, generator=g)
    outs = {}
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        aa = a.detach().clone().to(dt).requires_grad_(True)
        bb = b.detach().clone().to(dt).requires_grad_(True)
        out = ref_losses.CustomCLIPLoss(temperature)(aa, bb)
        out["total_loss"].backward()
        outs[tag] = {"clip_loss": out["clip_loss"].detach().clone(), "da": aa.grad.clone(), "db": bb.grad.clone()}
    torch.save(dict(name=name, a=a, b=b, temperature=temperature, **outs), os.path.join(HERE, f"clip_{name}.pt"))
    print("clip", name, float(outs["f32"]["clip_loss"]))


def pairwise_case(ref_losses, name, B, D, s, seed):
    """SPARCLoss.pairwise_contrastive_loss alone (losses.py:145-163)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, D, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(B, D, generator=g, dtype=torch.float64).requires_grad_(True)
    mod = ref_losses.SPARCLoss(cfg(0.5, 1.0, 1.0, s))
    loss = mod.pairwise_contrastive_loss(a, b)
    loss.backward()
    torch.save(dict(name=name, a=a.detach().float(), b=b.detach().float(), s=s, loss=loss.detach(),
                    da=a.grad.clone(), db=b.grad.clone()), os.path.join(HERE, f"pairwise_{name}.pt"))
    print("pairwise", name, float(loss))


def masked_pairwise_case(ref_losses, name, B, T, D, s, seed):
    """SPARCLoss.masked_pairwise_contrastive_loss alone (losses.py:165-197), all-True mask (the only case where the
    reference is finite) plus the per-sample-truncated evaluation of a padded mask (pins the 'truncate' semantics)."""
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, T, D, generator=g, dtype=torch.float64).requires_grad_(True)
    b = torch.randn(B, T, D, generator=g, dtype=torch.float64).requires_grad_(True)
    mod = ref_losses.SPARCLoss(cfg(0.5, 1.0, 1.0, s))
    full = torch.ones(B, T, dtype=torch.bool)
    loss = mod.masked_pairwise_contrastive_loss(a, b, full)
    loss.backward()
    out = dict(name=name, a=a.detach().float(), b=b.detach().float(), s=s, loss=loss.detach(), da=a.grad.clone(),
               db=b.grad.clone())
    # padded mask, truncated per sample: sum of per-sample (mean CE * count) / total count
    lens = torch.randint(2, T + 1, (B,), generator=g)
    lens[0] = T
    mask = torch.arange(T)[None, :] < lens[:, None]
    a2 = a.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    tot = 0.0
    for i in range(B):
        n = int(lens[i])
        tot = tot + mod.masked_pairwise_contrastive_loss(a2[i:i + 1, :n], b2[i:i + 1, :n], torch.ones(1, n, dtype=torch.bool)) * n
    n_valid = float(torch.tensor(int(lens.sum())) + 1e-8)
    lt = tot / n_valid
    lt.backward()
    out.update(mask=mask, loss_trunc=lt.detach(), da_trunc=a2.grad.clone(), db_trunc=b2.grad.clone(),
               ref_padded_is_nan=bool(torch.isnan(mod.masked_pairwise_contrastive_loss(a.detach(), b.detach(), mask))))
    torch.save(out, os.path.join(HERE, f"maskedpair_{name}.pt"))
    print("masked pairwise", name, float(loss), float(lt), "ref padded NaN:", out["ref_padded_is_nan"])


def adamspd_case(ref_opt, name, sizes, steps, lr, betas, eps, wd, amsgrad, seed, with_pre=True, none_grad_idx=()):
    g = torch.Generator().manual_seed(seed)
    p0 = [torch.randn(*s, generator=g) * 0.02 for s in sizes]
    pre = [p + 1e-3 * torch.randn(*p.shape, generator=g) for p in p0] if with_pre else None
    grads = [[torch.randn(*s, generator=g) * 1e-3 for s in sizes] for _ in range(steps)]
    params = [torch.nn.Parameter(p.clone()) for p in p0]
    opt = ref_opt.AdamSPD([{"params": params, "pre": pre}], lr=lr, betas=betas, eps=eps, weight_decay=wd,
                          amsgrad=amsgrad)
    snaps = {}
    for t in range(steps):
        for j, p in enumerate(params):
            p.grad = None if j in none_grad_idx and t % 2 == 1 else grads[t][j].clone()
        opt.step()
        if t + 1 in (1, 2, steps // 2, steps):
            snaps[t + 1] = [p.detach().clone() for p in params]
    state = [{k: (x.clone() if torch.is_tensor(x) else x) for k, x in opt.state[p].items() if k != "hyper"}
             for p in params]
    torch.save(dict(name=name, sizes=sizes, steps=steps, lr=lr, betas=betas, eps=eps, wd=wd, amsgrad=amsgrad,
                    p0=p0, pre=pre, grads=grads, snaps=snaps, state=state, none_grad_idx=tuple(none_grad_idx)),
               os.path.join(HERE, f"adamspd_{name}.pt"))
    print("adamspd", name, "steps", steps, "p[0][:3]", params[0].detach().flatten()[:3].tolist())


def statedict_case(ref_opt):
    """optimizer.state_dict() of the reference AdamSPD after 3 steps (checkpoint compatibility, finetuner.py:256-273)."""
    g = torch.Generator().manual_seed(41)
    sizes = [(5,), (3, 4)]
    params = [torch.nn.Parameter(torch.randn(*s, generator=g)) for s in sizes]
    pre = [p.detach().clone() + 0.01 for p in params]
    opt = ref_opt.AdamSPD([{"params": params, "pre": pre}], lr=1e-3, weight_decay=0.1, amsgrad=True)
    for _ in range(3):
        for p in params:
            p.grad = torch.randn(*p.shape, generator=g)
        opt.step()
    torch.save(dict(state_dict=opt.state_dict(), params=[p.detach().clone() for p in params], sizes=sizes),
               os.path.join(HERE, "statedict_ref.pt"))
    print("state_dict keys", sorted(opt.state_dict()["state"][0].keys()), sorted(opt.state_dict()["param_groups"][0].keys()))


def count_cases(ref_losses):
    """CountLoss (losses.py:267-309) and CLIPCountLoss (:39-133) on seeded inputs, fp64 and fp32, autograd gradients."""
    g = torch.Generator().manual_seed(31)
    B, C, D, T, alpha = 6, 5, 32, 0.07, 0.7
    out = dict(B=B, C=C, D=D, T=T, alpha=alpha)
    base = dict(la=torch.randn(B, B, generator=g, dtype=torch.float64) * 3, lb=torch.randn(B, B, generator=g, dtype=torch.float64) * 3,
                ei=torch.randn(B, D, generator=g, dtype=torch.float64), ek=torch.randn(B, D, generator=g, dtype=torch.float64),
                cf=torch.randn(B, C, D, generator=g, dtype=torch.float64))
    out["inputs"] = base
    for dt in (torch.float64, torch.float32):
        xs = {k: v.to(dt).clone().requires_grad_(True) for k, v in base.items()}
        res = ref_losses.CountLoss(T, alpha)(xs["la"], xs["lb"], xs["ei"], xs["ek"], xs["cf"])
        res["total_loss"].backward()
        out[str(dt)] = dict(losses={k: v.detach().clone() for k, v in res.items()}, grads={k: v.grad.clone() for k, v in xs.items()})
    torch.save(out, os.path.join(HERE, "countloss_b6_c5_d32.pt"))
    print("countloss", {k: float(v) for k, v in out["torch.float64"]["losses"].items()})
    Bi, nt, D2 = 4, 3, 24
    img = torch.randn(Bi, D2, generator=g, dtype=torch.float64)
    txt = torch.randn(Bi * nt, D2, generator=g, dtype=torch.float64)
    out2 = dict(B=Bi, nt=nt, D=D2, T=0.07, img=img, txt=txt)
    for dt in (torch.float64, torch.float32):
        a = img.to(dt).clone().requires_grad_(True); b = txt.to(dt).clone().requires_grad_(True)
        res = ref_losses.CLIPCountLoss(0.07, 0.5)(a, b, torch.arange(Bi * nt))
        res["total_loss"].backward()
        out2[str(dt)] = dict(losses={k: v.detach().clone() for k, v in res.items()}, da=a.grad.clone(), db=b.grad.clone())
    torch.save(out2, os.path.join(HERE, "clipcount_b4_t3_d24.pt"))
    print("clipcount", {k: (float(v), str(v.dtype)) for k, v in out2["torch.float32"]["losses"].items()})


def adamspd_amp_case(ref_opt, name, sizes, steps, lr, wd, max_norm, seed, inf_steps=(5, 11)):
    """The reference's update sequence under AMP (finetune/finetuner.py:147-154): scaler.unscale_ -> clip_grad_norm_ ->
    scaler.step -> scaler.update, with the unmodified AdamSPD and torch's GradScaler on CPU.  Gradients are handed over
    already multiplied by the current scale (what backward of the scaled loss produces); on `inf_steps` one gradient
    holds an inf, so the step is skipped and the scale backs off."""
    g = torch.Generator().manual_seed(seed)
    p0 = [torch.randn(*s, generator=g) * 0.02 for s in sizes]
    pre = [p + 1e-3 * torch.randn(*p.shape, generator=g) for p in p0]
    mags = [float(0.5 + 3.0 * torch.rand(1, generator=g)) for _ in range(steps)]       # some steps clip, some do not
    grads = [[torch.randn(*s, generator=g) * 1e-3 * mags[t] for s in sizes] for t in range(steps)]
    params = [torch.nn.Parameter(p.clone()) for p in p0]
    opt = ref_opt.AdamSPD([{"params": params, "pre": pre}], lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    scaler = torch.amp.GradScaler("cpu", init_scale=1024.0, growth_interval=3)
    scaler.scale(torch.zeros(1))                                   # lazy init of the scale tensor
    snaps, norms, scales, skipped = {}, [], [], []
    for t in range(steps):
        scale = float(scaler.get_scale())
        scales.append(scale)
        for j, p in enumerate(params):
            p.grad = grads[t][j] * scale
        if t in inf_steps:
            params[2].grad.view(-1)[3] = float("inf")
        scaler.unscale_(opt)
        norms.append(float(torch.nn.utils.clip_grad_norm_(params, max_norm)))
        before = params[0].detach().clone()
        scaler.step(opt)
        scaler.update()
        skipped.append(bool(torch.equal(before, params[0].detach())))
        if t + 1 in (1, 2, 6, 7, 12, steps):
            snaps[t + 1] = [p.detach().clone() for p in params]
    state_steps = [opt.state[p]["step"] for p in params]
    torch.save(dict(name=name, sizes=sizes, steps=steps, lr=lr, wd=wd, max_norm=max_norm, p0=p0, pre=pre, grads=grads,
                    inf_steps=tuple(inf_steps), snaps=snaps, norms=norms, scales=scales, skipped=skipped,
                    state_steps=state_steps, final_scale=float(scaler.get_scale())),
               os.path.join(HERE, f"ampstep_{name}.pt"))
    print("adamspd amp", name, "clipped steps", sum(n > max_norm for n in norms if n == n and n != float("inf")),
          "skipped", sum(skipped), "final scale", float(scaler.get_scale()), "steps", state_steps[0])


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    ref_losses, ref_opt = import_reference()
    if "--only-statedict" in sys.argv:
        statedict_case(ref_opt)
        return
    if "--only-count" in sys.argv:
        count_cases(ref_losses)
        return
    if "--only-adamspd-amp" in sys.argv:
        adamspd_amp_case(ref_opt, "s16", [(1,), (7,), (33, 31), (4099,), (64, 64), (8193,)], 16, 2e-5, 0.1, 0.2, seed=21)
        return
    if "--only-masked-pairwise" in sys.argv:
        masked_pairwise_case(ref_losses, "b3_t20_d48", 3, 20, 48, 3.0, seed=12)
        return
    # SPARC: thr = 1/P is well conditioned (SURVEY finding 2); one trainer-default case (thr=.5, s=.07)
    sparc_case(ref_losses, "b4_p50_d64_thrP", 4, 50, 77, 64, 1.0 / 50, 1.0, 1.0, 1.0, seed=1)
    sparc_case(ref_losses, "b3_p197_d32_thrP_w", 3, 197, 77, 32, 1.0 / 197, 0.7, 1.3, 2.5, seed=2)
    sparc_case(ref_losses, "b2_p50_d64_thr05_s007", 2, 50, 77, 64, 0.5, 1.0, 1.0, 0.07, seed=3)
    sparc_case(ref_losses, "b2_p33_t20_d48_local", 2, 33, 20, 48, 1.0 / 33, 1.0, 1.0, 3.0, seed=4,
               backprop="loss_vl_local")
    sparc_masked_case(ref_losses, "b4_p50_d64", 4, 50, 77, 64, 1.0 / 50, 1.5, seed=5)
    clip_case(ref_losses, "b16_d64", 16, 64, 0.07, seed=6)
    clip_case(ref_losses, "b5_d40_t1", 5, 40, 1.0, seed=7)
    pairwise_case(ref_losses, "b12_d32", 12, 32, 4.0, seed=8)
    masked_pairwise_case(ref_losses, "b3_t20_d48", 3, 20, 48, 3.0, seed=12)
    count_cases(ref_losses)
    statedict_case(ref_opt)
    sizes = [(1,), (7,), (33, 31), (4099,), (64, 64)]
    adamspd_case(ref_opt, "s20", sizes, 20, 2e-5, (0.9, 0.999), 1e-8, 0.1, False, seed=9)
    adamspd_case(ref_opt, "s12_ams_lr1e3", sizes, 12, 1e-3, (0.9, 0.98), 5e-6, 0.2, True, seed=10)
    adamspd_case(ref_opt, "s8_nopre_nonegrad", sizes, 8, 1e-3, (0.9, 0.999), 1e-8, 0.1, False, seed=11,
                 with_pre=False, none_grad_idx=(1, 3))
    adamspd_amp_case(ref_opt, "s16", [(1,), (7,), (33, 31), (4099,), (64, 64), (8193,)], 16, 2e-5, 0.1, 0.2, seed=21)


if __name__ == "__main__":
    main()
