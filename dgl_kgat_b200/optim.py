"""Adam for the KGAT parameters as ONE library launch per step.

``FusedAdam(params, lr)`` is a drop-in for ``torch.optim.Adam(params, lr)`` as the reference uses it
(kgat.py:101): same update rule, same ``state_dict`` keys (``step``, ``exp_avg``, ``exp_avg_sq``), so
checkpoints move between the two.  Parameters must be dense fp32 CUDA tensors; anything else raises.
"""
from __future__ import annotations

import ctypes

import torch as th

from . import _lib

__all__ = ["FusedAdam"]


class FusedAdam(th.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    def _launch(self, group, ps, grads, step):
        """One library call per 16 tensors; the pointer tables of everything but the gradients are cached
        (parameters and moments never move), so a step costs a few microseconds of host time."""
        lib = _lib.load()
        key = tuple(id(p) for p in ps)
        plan = self._plans.get(key) if hasattr(self, "_plans") else None
        if plan is None:
            if not hasattr(self, "_plans"):
                self._plans = {}
            plan = []
            for lo in range(0, len(ps), 16):
                sub = ps[lo: lo + 16]
                n = len(sub)
                arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])
                plan.append((lo, n, arr(sub), arr([self.state[p]["exp_avg"] for p in sub]),
                             arr([self.state[p]["exp_avg_sq"] for p in sub]),
                             (ctypes.c_int64 * n)(*[p.numel() for p in sub]),
                             tuple(self.state[p]["exp_avg"].data_ptr() for p in sub)))
            self._plans[key] = plan
        b1, b2 = group["betas"]
        with th.cuda.device(ps[0].device):
            st = _lib.stream_ptr()
            for lo, n, a_p, a_m, a_v, a_n, m_ptrs in plan:
                a_g = (ctypes.c_void_p * n)(*[g.data_ptr() for g in grads[lo: lo + n]])
                _lib.check(lib.kgat_adam_step(n, a_p, a_g, a_m, a_v, a_n, group["lr"], b1, b2, group["eps"],
                                              group["weight_decay"], step, st), "kgat_adam_step")

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans, self._fast = {}, {}                  # the moment tensors were replaced

    @th.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with th.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            # steady state: the same parameters as last time, all on the same step count
            fast = self._fast.get(gi) if hasattr(self, "_fast") else None
            if fast is not None and len(fast[0]) == len(ps) and all(a is b for a, b in zip(fast[0], ps)) \
                    and all(p.grad.is_contiguous() for p in ps):
                step = fast[1] + 1
                self._launch(group, ps, [p.grad for p in ps], step)
                for p in ps:
                    self.state[p]["step"] = step
                self._fast[gi] = (ps, step)
                continue
            fresh = False
            for p in ps:
                st = self.state[p]
                if not st:
                    if not (p.is_cuda and p.dtype == th.float32 and p.is_contiguous()) or p.grad.is_sparse:
                        raise _lib.KgatError("FusedAdam needs dense contiguous fp32 CUDA parameters")
                    st["step"] = 0
                    st["exp_avg"] = th.zeros_like(p, memory_format=th.preserve_format)
                    st["exp_avg_sq"] = th.zeros_like(p, memory_format=th.preserve_format)
                    fresh = True
            if fresh and hasattr(self, "_plans"):
                self._plans = {}
            grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
            by_step = {}
            for i, p in enumerate(ps):                    # parameters that joined later keep their own count
                by_step.setdefault(int(self.state[p]["step"]), []).append(i)
            for s, idx in by_step.items():
                self._launch(group, [ps[i] for i in idx], [grads[i] for i in idx], s + 1)
            for p in ps:
                self.state[p]["step"] = int(self.state[p]["step"]) + 1
            if not hasattr(self, "_fast"):
                self._fast = {}
            if len(by_step) == 1:
                self._fast[gi] = (ps, next(iter(by_step)) + 1)
            else:
                self._fast.pop(gi, None)
        return loss
