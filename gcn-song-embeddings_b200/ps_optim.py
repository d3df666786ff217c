"""Adam for the PinSage trainer as ONE fused kernel per step (ps_adam_step) over flat fp32 buffers, behind
torch.optim.Adam's interface: `param_groups` (the ExponentialLR scheduler of the reference drives `lr` there,
pinsage_training.py:147-148), `state_dict()` / `load_state_dict()` in torch's own format, so a `state.pt` written by the
reference loads here and vice versa (pinsage_training.py:277-295).

torch.optim.Adam's foreach path issues ~10 multi-tensor kernels per step; at the reference's default sizes (a 0.5 M
parameter model, 384-node batches) they are a quarter of the step.  Parameters, their gradients (ps_engine.Engine's
flat_grad) and both moment buffers are contiguous, so the whole update is one pass over 4 x 2.3 MB.
"""
from __future__ import annotations

import torch

import ps_native


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, engine=None):
        params = list(params)
        super().__init__(params, lr=lr, betas=betas, eps=eps)
        self._engine = engine
        self._params = params
        total = sum(p.numel() for p in params)
        dev = params[0].device
        self._flat_p = torch.empty(total, dtype=torch.float32, device=dev)
        self._flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        self._flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        self._t = 0
        self.grad_scale = 1.0  # data parallel: 1 / world_size folded into the update (the allreduce sums)
        off = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                view = self._flat_p[off: off + n].view_as(p)
                view.copy_(p.data)
                p.data = view  # the parameter now lives in the flat buffer (in-place updates elsewhere keep working)
                off += n
        self._bind_state()

    def _bind_state(self):
        off = 0
        for p in self._params:
            n = p.numel()
            self.state[p] = {"step": torch.tensor(float(self._t)),
                             "exp_avg": self._flat_m[off: off + n].view_as(p),
                             "exp_avg_sq": self._flat_v[off: off + n].view_as(p)}
            off += n

    def _flat_grad(self):
        eng = self._engine
        flat = getattr(eng, "flat_grad", None) if eng is not None else None
        p0 = self._params[0]
        if flat is not None and p0.grad is not None and p0.grad.data_ptr() == flat.data_ptr() and flat.numel() == self._flat_p.numel():
            return flat  # the engine's gradients are views of one buffer in parameter order
        parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32) for p in self._params]
        return torch.cat(parts)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        if g.get("weight_decay", 0) or g.get("amsgrad", False) or g.get("maximize", False):
            raise NotImplementedError("FlatAdam covers the reference's optimiser: Adam(lr) with torch defaults")
        p0 = self._params[0]
        if p0.data_ptr() != self._flat_p.data_ptr():  # someone re-pointed .data (e.g. model.to()): adopt the new values
            off = 0
            for p in self._params:
                n = p.numel()
                view = self._flat_p[off: off + n].view_as(p)
                view.copy_(p.data); p.data = view
                off += n
        self._t += 1
        b1, b2 = g["betas"]
        ps_native.adam_step(self._flat_p, self._flat_grad(), self._flat_m, self._flat_v, float(g["lr"]), b1, b2, g["eps"], self._t,
                            grad_scale=self.grad_scale)
        return loss

    def state_dict(self):
        for st in self.state.values():
            st["step"] = torch.tensor(float(self._t))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = [float(st["step"]) for st in self.state.values() if "step" in st]
        self._t = int(max(steps)) if steps else 0
        off = 0
        for p in self._params:  # loaded moments -> the flat buffers
            n = p.numel()
            st = self.state.get(p, {})
            if "exp_avg" in st:
                self._flat_m[off: off + n].view_as(p).copy_(st["exp_avg"])
                self._flat_v[off: off + n].view_as(p).copy_(st["exp_avg_sq"])
            else:
                self._flat_m[off: off + n].zero_(); self._flat_v[off: off + n].zero_()
            off += n
        self._bind_state()
