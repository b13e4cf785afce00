"""The optimiser the reference builds for `cs` (utils/utils.py:112-141: torch.optim.Adam, encoder at lr / 10) with its step -
model.py:121 `self.optimizer.step()` - as ONE launch of `pp_adam_step_multi` over every parameter tensor.

`FusedAdam` IS a `torch.optim.Adam` (capturable: device step counters and tensor learning rates, usable inside a captured CUDA
graph): same constructor groups, same `state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), same schedulers.
Only `step()` differs: torch's multi-tensor kernel (six launches for RN50-DeepLabv3+, 2.7 TB/s in the step's profile) is
replaced by one kernel that takes the whole tensor list in its parameter space.  Anything the kernel does not cover (amsgrad,
maximize, a parameter without a gradient, non-fp32 parameters, CPU tensors) goes through `torch.optim.Adam.step` unchanged.
"""
import torch

from . import _lib


class FusedAdam(torch.optim.Adam):
    def __init__(self, param_groups):
        super().__init__(param_groups, fused=True, capturable=True)
        self._steps = None  # fp32 [n]: every parameter's `state["step"]` is a view of one element -> one add per step
        self._plan = None

    def _ours(self):
        if len(self.param_groups) > 8:
            return False
        for g in self.param_groups:
            if g.get("amsgrad") or g.get("maximize") or g.get("differentiable") or not torch.is_tensor(g["lr"]):
                return False
            for p in g["params"]:
                if p.grad is None or p.grad.is_sparse or p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() \
                        or not p.grad.is_contiguous():
                    return False
        return True

    def _bind_state(self, params):
        """moments and step counters in torch's own state layout; the step counters of all parameters share one buffer"""
        dev = params[0].device
        n = len(params)
        if self._steps is None or self._steps.numel() != n or self._steps.device != dev:
            self._steps = torch.zeros(n, dtype=torch.float32, device=dev)
        base = self._steps.data_ptr()
        stale = [i for i, p in enumerate(params) if "step" not in self.state[p] or self.state[p]["step"].data_ptr() != base + 4 * i]
        for i in stale:  # first step, or load_state_dict() replaced the tensors: move the values into the shared buffer
            st = self.state[params[i]]
            if "step" in st:
                self._steps[i] = torch.as_tensor(st["step"], dtype=torch.float32).to(dev)
            st["step"] = self._steps[i]
            for k in ("exp_avg", "exp_avg_sq"):
                if k not in st:
                    st[k] = torch.zeros_like(params[i], memory_format=torch.preserve_format)

    @torch.no_grad()
    def step(self, closure=None):
        if not self._ours():
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        params = [p for g in self.param_groups for p in g["params"]]
        self._bind_state(params)
        grads = [p.grad for p in params]
        m = [self.state[p]["exp_avg"] for p in params]
        v = [self.state[p]["exp_avg_sq"] for p in params]
        lrs = [g["lr"] for g in self.param_groups]
        key = _lib.AdamPlan.make_key(params, grads, m, v, lrs, self._steps)
        hyper = [(g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"]) for g in self.param_groups]
        if self._plan is None or self._plan.key != key or self._plan.hyper != hyper:
            group = [gi for gi, g in enumerate(self.param_groups) for _ in g["params"]]
            self._plan = _lib.AdamPlan(params, grads, m, v, group, lrs, *zip(*hyper), self._steps)
            self._plan.hyper = hyper
        self._steps.add_(1)  # torch increments before it forms the bias corrections
        self._plan.launch()
        return loss
