"""CPU oracle for AdamSPD.  TEST INFRASTRUCTURE ONLY (see oracle/losses_oracle.py header).

Restates ``finetune/optimizers.py:100-157`` (AdamSPD.adam + _ratio) as a pure
function over plain tensors, in the reference's evaluation order:
``(step_size * m) / denom``, ``sqrt(v) / sqrt(bc2) + eps``, bias corrections as
Python doubles.  Pinned by ``tests/golden/adamspd_*.pt`` (made by running the
reference optimizer itself, ``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch


def adamspd_tensor_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                        pre: Optional[torch.Tensor], step: int, lr: float, beta1: float, beta2: float,
                        eps: float, weight_decay: float, vmax: Optional[torch.Tensor] = None):
    """One tensor, one step, in place on p, m, v (and vmax).  Returns (projected: bool, ratio: float)."""
    bc1 = 1 - beta1 ** step                                   # optimizers.py:123
    bc2 = 1 - beta2 ** step                                   # optimizers.py:124
    m.mul_(beta1).add_(g, alpha=1 - beta1)                    # :128
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)             # :129
    if vmax is not None:                                      # :131-135
        torch.maximum(vmax, v, out=vmax)
        denom = (vmax.sqrt() / math.sqrt(bc2)).add_(eps)
    else:                                                     # :137
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    step_size = lr / bc1                                      # :139
    new_p = p - step_size * m / denom                         # :142-143
    pre_t = pre if pre is not None else torch.zeros_like(p)   # :146
    cond = -torch.sum(g * (p - pre_t))                        # :147
    projected, ratio = False, 0.0
    if cond < 0.0:                                            # :148
        curr, prev = torch.norm(new_p - pre_t), torch.norm(p - pre_t)      # :155
        r = torch.nn.functional.hardtanh((curr - prev) / curr, 0.0, 1.0)   # :156-157
        new_p = new_p - weight_decay * r * (new_p - pre_t)    # :150
        projected, ratio = True, float(r)
    p.copy_(new_p)                                            # :151
    return projected, ratio


def adamspd_step(params: List[torch.Tensor], grads: List[Optional[torch.Tensor]], ms, vs, pres, steps: List[int],
                 lr, betas, eps, weight_decay, vmaxs=None):
    """Whole-list step (optimizers.py:31-98 + :100-152): params with grad None are skipped,
    their step counter does not advance."""
    stats = []
    for j, p in enumerate(params):
        if grads[j] is None:
            stats.append(None)
            continue
        steps[j] += 1                                          # :81
        stats.append(adamspd_tensor_step(p, grads[j], ms[j], vs[j], None if pres is None else pres[j],
                                         steps[j], lr, betas[0], betas[1], eps, weight_decay,
                                         None if vmaxs is None else vmaxs[j]))
    return stats


# --------------------------------------------------------------------------
# AMP prologue folded into the step:
#   scaler.unscale_(optimizer); clip_grad_norm_(params, max_norm); scaler.step(optimizer)
# finetune/finetuner.py:150-152.  The arithmetic is torch's (torch/amp/grad_scaler.py,
# torch/nn/utils/clip_grad.py, version of this image: 2.11); pinned by
# tests/golden/ampstep_*.pt (made with the reference AdamSPD + torch.amp.GradScaler on CPU).
# --------------------------------------------------------------------------
def amp_unscale_clip(grads: List[Optional[torch.Tensor]], scale: float, max_norm: Optional[float]):
    """Returns (effective grads or None when the step must be skipped, total_norm, found_inf).
    unscale_: g *= (1/scale as double -> float), found_inf if any raw g is non-finite;
    clip_grad_norm_: per-tensor fp32 2-norms, norm of their stack, g *= clamp(max_norm / (total + 1e-6), max=1)."""
    inv = torch.tensor(scale, dtype=torch.float32).double().reciprocal().float()
    present = [g for g in grads if g is not None]
    found_inf = any(not bool(torch.isfinite(g).all()) for g in present)
    un = [None if g is None else (g if float(inv) == 1.0 else g * inv) for g in grads]
    norms = [torch.linalg.vector_norm(g, 2.0) for g in un if g is not None]
    total = torch.linalg.vector_norm(torch.stack(norms), 2.0)
    if max_norm is not None:
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        un = [None if g is None else g * coef for g in un]
    return (None if found_inf else un), float(total), found_inf


def amp_scale_update(scale: float, growth_tracker: int, found_inf: bool, growth_factor: float = 2.0,
                     backoff_factor: float = 0.5, growth_interval: int = 2000):
    """GradScaler.update (torch._amp_update_scale_): back off on inf, grow after `growth_interval` clean steps."""
    if found_inf:
        return scale * backoff_factor, 0
    growth_tracker += 1
    if growth_tracker == growth_interval:
        return scale * growth_factor, 0
    return scale, growth_tracker
