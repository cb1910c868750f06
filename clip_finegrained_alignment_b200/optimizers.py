"""AdamSPD as a torch.optim.Optimizer backed by one multi-tensor sm_100a kernel pair.

Drop-in for the reference's `finetune/optimizers.py:8-157`: same constructor, same validation and
error behaviour, same param-group convention (`{'params': [...], 'pre': [...] or None}`), same
`state_dict()` layout (`step` as a Python int, `exp_avg`, `exp_avg_sq`, unused `hyper`,
`max_exp_avg_sq` when amsgrad).  The per-tensor Python loop with ~20 launches and one host sync per
tensor (optimizers.py:113-152) is replaced by `cfa_adamspd_step` (include/cfa_b200.h): no host sync.
"""
from __future__ import annotations

import math

import numpy as np
import torch
from torch.optim.optimizer import Optimizer

from . import _lib

_TENSOR_DT = np.dtype([
    ("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("pre", "<u8"), ("vmax", "<u8"), ("numel", "<i8"),
    ("beta1", "<f4"), ("omb1", "<f4"), ("beta2", "<f4"), ("omb2", "<f4"), ("eps", "<f4"),
    ("step_size", "<f4"), ("sqrt_bc2", "<f4"), ("wd", "<f4")])
assert _TENSOR_DT.itemsize == 88


class _Plan:
    """Cached launch plan for one pattern of (param has grad) over all param groups.  The static half of the
    descriptor table (p, m, v, pre, vmax pointers, sizes) and the chunk table are built once; per step only the
    gradient pointers and a handful of per-(group, step) scalars are refreshed."""

    def __init__(self, opt, all_params, has_grad, amsgrad, device):
        self.all_params = all_params
        self.has_grad = has_grad
        self.amsgrad = amsgrad
        entries, combo_key, combo_index = [], {}, []
        k = 0
        for gi, group in enumerate(opt.param_groups):
            pre_list = group["pre"]
            for j, p in enumerate(group["params"]):
                if has_grad[k]:
                    st = opt.state[p]
                    pre = pre_list[j] if pre_list is not None else None
                    entries.append((p, st, pre))
                    key = (gi, st["step"])                      # params of one group normally share the step count
                    combo_index.append(combo_key.setdefault(key, len(combo_key)))
                k += 1
        self.entries = entries
        self.states = [e[1] for e in entries]
        self.grad_params = [e[0] for e in entries]
        self.combos = list(combo_key.keys())                     # (group index, step count at plan time)
        self.combo_index = np.asarray(combo_index, dtype=np.int64)
        self.steps_since_build = 0
        n = self.n = len(entries)
        chunk = _lib.lib.cfa_adamspd_chunk_elems()
        tab = np.zeros(n, dtype=_TENSOR_DT)
        chunks = []
        for i, (p, st, pre) in enumerate(entries):
            tab["p"][i] = p.data_ptr()
            tab["m"][i] = st["exp_avg"].data_ptr()
            tab["v"][i] = st["exp_avg_sq"].data_ptr()
            tab["pre"][i] = 0 if pre is None else pre.data_ptr()
            tab["vmax"][i] = st["max_exp_avg_sq"].data_ptr() if amsgrad else 0
            tab["numel"][i] = p.numel()
            nc = (p.numel() + chunk - 1) // chunk
            chunks.append(np.stack([np.full(nc, i, dtype=np.int32), np.arange(nc, dtype=np.int32)], axis=1))
        self.static = tab
        self.static_key = (tab["p"].tobytes(), tab["m"].tobytes(), tab["pre"].tobytes())
        ch = np.concatenate(chunks, axis=0) if chunks else np.zeros((0, 2), np.int32)
        self.n_chunks = int(ch.shape[0])
        self.d_chunks = torch.from_numpy(np.ascontiguousarray(ch)).to(device)
        # two pinned staging buffers so that refilling never races the previous step's async copy
        self.h_tab = [torch.empty(n * _TENSOR_DT.itemsize, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.h_np = [t.numpy().view(_TENSOR_DT) for t in self.h_tab]
        for a in self.h_np:
            a[:] = tab
        self.h_evt = [torch.cuda.Event() for _ in range(2)]
        self.h_used = [False, False]
        self.flip = 0
        self.d_tab = torch.empty(n * _TENSOR_DT.itemsize, dtype=torch.uint8, device=device)
        self.d_reduce = torch.empty(4 * n, dtype=torch.float64, device=device)      # 3 SPD sums + (amp) grad sum of squares
        self.h_flag = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.d_amp = torch.zeros(4, dtype=torch.float32, device=device)             # inv_scale, clip_coef, total_norm, found_inf
        self.d_stats = torch.zeros(n, 2, dtype=torch.float32, device=device)


class AdamSPD(Optimizer):
    """Adam with Selective Projection Decay (https://arxiv.org/abs/2411.01713), reference API."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        # validation mirrors optimizers.py:11-20
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        if not 0.0 <= weight_decay:
            raise ValueError("Invalid weight_decay value: {}".format(weight_decay))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        self._plan = None
        self._pending_skip = None       # (event, pinned found_inf) of the last amp_step: settles the host step counters

    def __setstate__(self, state):
        super().__setstate__(state)
        for group in self.param_groups:
            group.setdefault("amsgrad", False)
        self._plan = None

    # ------------------------------------------------------------------
    @property
    def last_step_stats(self):
        """[n_tensors, 2] device tensor of (projected?, ratio) from the most recent step, in the order
        params-with-grad were visited.  Reading it on the host synchronises; step() itself never does."""
        return None if self._plan is None else self._plan.d_stats

    def _settle_skipped_step(self):
        """amp_step decides on the DEVICE whether the step is skipped (non-finite gradients); the host-side `step`
        counters (Python ints, as in the reference state_dict) are corrected here, one call later, when the flag has
        long arrived in pinned memory — no stall on the step path."""
        pend = self._pending_skip
        if pend is None:
            return
        self._pending_skip = None
        evt, flag, plan = pend
        evt.synchronize()
        if float(flag[0]) != 0.0:
            for st in plan.states:
                st["step"] -= 1
            plan.steps_since_build -= 1

    def state_dict(self):
        self._settle_skipped_step()
        return super().state_dict()

    @torch.no_grad()
    def amp_step(self, grad_scaler=None, max_grad_norm=None, closure=None):
        """`scaler.unscale_(opt); clip_grad_norm_(params, max_grad_norm); scaler.step(opt)` (finetuner.py:150-152) in
        one call: the unscale and clip passes over the gradients are folded into the multi-tensor step
        (cfa_adamspd_step_amp: one extra READ of g instead of two read-modify-write passes, no host sync; the
        reference's GradScaler.step synchronises on found_inf every step).  Gradients stay as they are in memory
        (still scaled) — the reference zeroes them right after (finetuner.py:154).  Returns the total gradient norm
        (0-dim device tensor), like clip_grad_norm_.  `scaler.update()` is called by the caller as before."""
        return self._step_impl(closure, amp=(grad_scaler, max_grad_norm))

    @torch.no_grad()
    def step(self, closure=None):
        return self._step_impl(closure, amp=None)

    def _step_impl(self, closure, amp):
        self._settle_skipped_step()
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()

        plan = self._plan
        if plan is None:
            all_params = [p for group in self.param_groups for p in group["params"]]
        else:
            all_params = plan.all_params
        grads_all = [p.grad for p in all_params]
        has_grad = tuple(g is not None for g in grads_all)
        if not any(has_grad):
            return loss if amp is None else torch.zeros(())
        grads = [g for g in grads_all if g is not None]
        if plan is None or plan.has_grad != has_grad or len(all_params) != sum(len(g["params"]) for g in self.param_groups):
            plan = self._build_plan(grads)
        for g in grads:
            if not g.is_contiguous() or g.dtype != torch.float32 or g.is_sparse:
                plan = self._build_plan(grads)           # slow path re-validates and raises the reference's errors
                grads = [g if g.is_contiguous() else g.contiguous() for g in grads]
                break
        for st in plan.states:                          # optimizers.py:81
            st["step"] += 1
        plan.steps_since_build += 1
        total_norm = self._launch(plan, grads, amp)
        return loss if amp is None else total_norm

    # ------------------------------------------------------------------
    def _build_plan(self, grads):
        """Slow path: (re)validate everything like the reference does, lazily create state, build the tables."""
        all_params, has_grad, amsgrad_all = [], [], None
        for group in self.param_groups:
            ams = bool(group["amsgrad"])
            any_grad = False
            for p in group["params"]:
                all_params.append(p)
                g = p.grad
                has_grad.append(g is not None)
                if g is None:
                    continue
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                any_grad = True
                state = self.state[p]
                if len(state) == 0:       # lazy init, optimizers.py:61-70
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["hyper"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if ams:
                        state["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if any_grad:
                group["pre"]              # KeyError('pre') like the reference (optimizers.py:146)
                if amsgrad_all is None:
                    amsgrad_all = ams
                elif amsgrad_all != ams:
                    raise _lib.CfaError("AdamSPD: mixing amsgrad and non-amsgrad groups in one optimizer is unsupported")
        k = 0
        tensors = []
        for group in self.param_groups:
            pre_list = group["pre"] if any(p.grad is not None for p in group["params"]) else None
            for j, p in enumerate(group["params"]):
                if p.grad is not None:
                    pre = pre_list[j] if pre_list is not None else None
                    st = self.state[p]
                    tensors += [p, p.grad, pre]
                    if p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                        raise _lib.CfaError("AdamSPD kernels are fp32 (the reference keeps fp32 master params); got "
                                            f"{p.dtype}/{p.grad.dtype}")
                    if not p.is_contiguous() or not st["exp_avg"].is_contiguous() or not st["exp_avg_sq"].is_contiguous():
                        raise _lib.CfaError("AdamSPD: parameters and optimizer state must be contiguous")
                    if pre is not None and (pre.shape != p.shape or not pre.is_contiguous() or pre.dtype != p.dtype
                                            or pre.device != p.device):
                        raise _lib.CfaError("AdamSPD: group['pre'][j] must match its parameter (shape, dtype, device, contiguous)")
                k += 1
        dev = _lib.require_cuda(*tensors)
        self._plan = _Plan(self, all_params, tuple(has_grad), bool(amsgrad_all), dev)
        return self._plan

    def _launch(self, plan, grads, amp=None):
        k = plan.flip
        plan.flip ^= 1
        if plan.h_used[k]:
            plan.h_evt[k].synchronize()      # normally long finished: two steps ago
        tab = plan.h_np[k]
        tab["g"] = np.fromiter((g.data_ptr() for g in grads), dtype=np.uint64, count=plan.n)
        # The reference re-reads group['pre'][j], group['amsgrad'] and the state tensors on every step (optimizers.py:58-98):
        # parameters may have been re-pointed (p.data = ...), the anchors replaced (re-anchoring SPD), a state tensor
        # reassigned or amsgrad toggled since the plan was built -> compare every cached pointer, rebuild on any change
        pptr = np.fromiter((p.data_ptr() for p in plan.grad_params), dtype=np.uint64, count=plan.n)
        stale = not np.array_equal(pptr, plan.static["p"])
        if not stale:
            ams = plan.amsgrad
            mptr = np.fromiter((st["exp_avg"].data_ptr() for st in plan.states), dtype=np.uint64, count=plan.n)
            vptr = np.fromiter((st["exp_avg_sq"].data_ptr() for st in plan.states), dtype=np.uint64, count=plan.n)
            stale = not (np.array_equal(mptr, plan.static["m"]) and np.array_equal(vptr, plan.static["v"]))
            if not stale:
                k2 = 0
                for gi, group in enumerate(self.param_groups):
                    if bool(group["amsgrad"]) != bool(ams):
                        stale = True
                        break
                    pre_list = group["pre"]
                    for j, p in enumerate(group["params"]):
                        if p.grad is None:
                            continue
                        pre = pre_list[j] if pre_list is not None else None
                        if (0 if pre is None else pre.data_ptr()) != int(plan.static["pre"][k2]):
                            stale = True
                            break
                        k2 += 1
                    if stale:
                        break
        if stale:
            self._plan = None
            plan = self._build_plan(grads)
            return self._launch(plan, grads, amp)
        # per-step scalars: Python doubles like the reference (optimizers.py:123-124,139), rounded once to fp32
        rows = np.empty((len(plan.combos), 8), dtype=np.float32)
        for ci, (gi, step0) in enumerate(plan.combos):
            group = self.param_groups[gi]
            beta1, beta2 = group["betas"]
            step = step0 + plan.steps_since_build
            bc1 = 1 - beta1 ** step
            bc2 = 1 - beta2 ** step
            rows[ci] = (beta1, 1 - beta1, beta2, 1 - beta2, group["eps"], group["lr"] / bc1, math.sqrt(bc2),
                        group["weight_decay"])
        cols = rows[plan.combo_index]
        for c, name in enumerate(("beta1", "omb1", "beta2", "omb2", "eps", "step_size", "sqrt_bc2", "wd")):
            tab[name] = cols[:, c]
        dev = plan.d_tab.device
        with torch.cuda.device(dev):
            plan.d_tab.copy_(plan.h_tab[k], non_blocking=True)
            plan.h_evt[k].record()
            plan.h_used[k] = True
            total_norm = None
            if amp is None:
                _lib.call("cfa_adamspd_step", plan.d_tab.data_ptr(), plan.n, plan.d_chunks.data_ptr(), plan.n_chunks,
                          plan.d_reduce.data_ptr(), plan.d_stats.data_ptr(), 0, int(plan.amsgrad), _lib.stream_ptr())
            else:
                scaler, max_norm = amp
                scale_ptr = 0
                enabled = scaler is not None and scaler.is_enabled()
                if enabled:
                    from torch.amp.grad_scaler import OptState
                    ost = scaler._per_optimizer_states[id(self)]
                    if ost["stage"] is OptState.UNSCALED:
                        raise RuntimeError("amp_step() replaces scaler.unscale_(): the gradients of this optimizer were "
                                           "already unscaled")
                    if ost["stage"] is OptState.STEPPED:
                        raise RuntimeError("amp_step() has already been called since the last update().")
                    if scaler._scale is None:
                        scaler._lazy_init_scale_growth_tracker(dev)
                    scale = scaler._scale
                    if scale.device != dev:
                        scale = scale.to(dev, non_blocking=True)
                    scale_ptr = scale.data_ptr()
                _lib.call("cfa_adamspd_step_amp", plan.d_tab.data_ptr(), plan.n, plan.d_chunks.data_ptr(), plan.n_chunks,
                          plan.d_reduce.data_ptr(), plan.d_stats.data_ptr(), scale_ptr,
                          float(max_norm) if max_norm is not None else 0.0, plan.d_amp.data_ptr(), 0, int(plan.amsgrad),
                          _lib.stream_ptr())
                res = plan.d_amp.clone()             # this step's {inv_scale, clip_coef, total_norm, found_inf}
                total_norm = res[2]
                if enabled:                          # what scaler.update() reads (torch/amp/grad_scaler.py)
                    ost["found_inf_per_device"] = {dev: res[3:4]}
                    ost["stage"] = OptState.STEPPED
                flag = plan.h_flag                   # settled at the start of the next call, so one buffer is enough
                flag.copy_(res[3:4], non_blocking=True)
                evt = torch.cuda.Event()
                evt.record()
                self._pending_skip = (evt, flag, plan)
        self._keepalive = grads      # contiguous copies (if any) must outlive the async launch
        return total_norm
