"""Fused optimizer step for the student: AdamW for every parameter and the global
gradient norm in one kernel launch.

Drop-in for `torch.optim.AdamW(student.parameters(), lr=args.lr, weight_decay=args.weight_decay)`
(scripts/phase5_big_run.py:1621) together with the gradient-norm loop at :1783-1789, which does one
`.item()` host synchronisation per parameter tensor.  State layout (`step`, `exp_avg`, `exp_avg_sq`) and
`state_dict()` are torch.optim.AdamW's own, so checkpoints (`save_checkpoint`, :1104-1125) load either way.
There is no CPU path: parameters must be fp32 CUDA tensors.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _ext


class FusedAdamW(torch.optim.AdamW):
    """torch.optim.AdamW with `step()` running as ONE multi-tensor CUDA kernel (+ a one-block reduction).

    After `step()`, `last_grad_norm` is a device scalar holding ||g||_2 over all parameters that had a
    gradient (what the reference logs as `total_norm`), available without a host sync.
    `grad_scale` multiplies the gradients first (pass 1/scaler.get_scale() under fp16 loss scaling)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self._plans = {}
        self.last_grad_norm: Optional[torch.Tensor] = None

    def _plan(self, gi, ps):
        """chunk table over (p, grad, exp_avg, exp_avg_sq); rebuilt only when a gradient tensor moved
        (zero_grad(set_to_none=True) re-allocates them every accumulation window)"""
        gkey = tuple(p.grad.data_ptr() for p in ps)
        hit = self._plans.get(gi)
        if hit is not None and hit[0] == gkey and hit[2] == len(ps):
            return hit[1]
        if hit is not None:
            _ext.lib().dinox_adamw_plan_destroy(hit[1])
        n = len(ps)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
        ne = (ctypes.c_int64 * n)(*[p.numel() for p in ps])
        h = ctypes.c_void_p()
        _ext.call("dinox_adamw_plan_create", arr(ps), arr([p.grad for p in ps]), arr([self.state[p]["exp_avg"] for p in ps]),
                  arr([self.state[p]["exp_avg_sq"] for p in ps]), ne, n, ctypes.byref(h))
        self._plans[gi] = (gkey, h, n)
        return h

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        norms = []
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize"):
                raise NotImplementedError("FusedAdamW: amsgrad / maximize are not supported")
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            fresh = [p for p in ps if len(self.state[p]) == 0]
            for p in fresh:
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise _ext.DinoxError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if fresh or gi not in self._plans:
                for p in ps:
                    if not (p.grad.is_cuda and p.grad.dtype == torch.float32 and p.grad.is_contiguous()):
                        raise _ext.DinoxError("FusedAdamW needs contiguous fp32 CUDA gradients")
                steps = {float(self.state[p]["step"]) for p in ps}
                if len(steps) != 1:
                    raise _ext.DinoxError("FusedAdamW: parameters of one group must share the step count")
            t = float(self.state[ps[0]]["step"]) + 1.0
            beta1, beta2 = group["betas"]
            h = self._plan(gi, ps)
            norm = torch.empty((), dtype=torch.float32, device=ps[0].device)
            _ext.call("dinox_adamw_step", h, float(group["lr"]), float(beta1), float(beta2), float(group["eps"]),
                      float(group["weight_decay"]), 1.0 - beta1 ** t, 1.0 - beta2 ** t, float(grad_scale),
                      ctypes.c_void_p(norm.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            torch._foreach_add_([self.state[p]["step"] for p in ps], 1.0)
            norms.append(norm)
        if norms:
            self.last_grad_norm = norms[0] if len(norms) == 1 else torch.stack(norms).square().sum().sqrt()
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        for _, h, _n in self._plans.values():   # moment tensors were replaced: rebuild the chunk tables
            _ext.lib().dinox_adamw_plan_destroy(h)
        self._plans = {}

    def __del__(self):
        try:
            for _, h, _n in self._plans.values():
                _ext.lib().dinox_adamw_plan_destroy(h)
        except Exception:
            pass
