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
