"""Fused Adam + EMA over the flat parameter arena.

Replaces ``optimizer.step()`` over 314 tensors plus the per-parameter Python
EMA loop of the reference trainer (trainers/ddpm_trainer.py:139-143,463-480)
with ONE launch of ``dmu_adam_ema`` (28 + 12 bytes per parameter of HBM
traffic).  Semantics are torch.optim.Adam's (bias-corrected, eps added after
the sqrt, optional L2 weight decay) followed by
``ema = decay * ema + (1 - decay) * param``.
"""

import torch

from . import _abi, ops


class FusedAdamEMA:
    def __init__(self, unet, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, ema_decay=None):
        self.unet = unet
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        self.step_count = 0
        self.m = self.v = self.ema = None
        self.step_dev = None      # device-resident copy of step_count: lets the update sit in a replayed CUDA graph

    def _ensure(self):
        eng = self.unet.engine
        if eng.flat is None:
            raise RuntimeError("FusedAdamEMA: run a forward pass first (the parameter arena is created lazily on the device)")
        if self.m is None or self.m.data_ptr() == 0 or self.m.numel() != eng.flat.numel() or self.m.device != eng.flat.device:
            self.m = torch.zeros_like(eng.flat)
            self.v = torch.zeros_like(eng.flat)
            self.ema = eng.flat.clone() if self.ema_decay is not None else None
        return eng

    def step(self, grad_scale: float = 1.0):
        """Apply one update using the gradient arena filled by the last backward."""
        self.begin_step()
        self.step_range(0, None, grad_scale)

    def begin_step(self, count_host: bool = True):
        """Once per optimisation step, before its step_range() calls: advances the step count on the host and on the device.
        count_host=False while the caller CAPTURES the step into a CUDA graph (the device increment is recorded, the host
        mirror is advanced by the caller once per replay)."""
        eng = self._ensure()
        if self.step_dev is None or self.step_dev.device != eng.flat.device:
            self.step_dev = torch.full((), self.step_count, device=eng.flat.device, dtype=torch.int64)
        self.step_dev.add_(1)
        if count_host:
            self.step_count += 1

    def step_range(self, lo: int = 0, hi=None, grad_scale: float = 1.0):
        """Update the arena elements [lo, hi) (16-byte aligned range starts) from the gradients of the same range.  The
        backward finishes its gradients part by part (Engine.run_backward's ``between``): the update of a finished range runs on
        another stream while the rest of the backward is still in flight; the bias corrections come from the device step count."""
        eng = self._ensure()
        hi = eng.flat.numel() if hi is None else hi
        if hi <= lo:
            return
        off = lambda t: t.data_ptr() + lo * 4
        _abi.check(_abi.lib().dmu_adam_ema(
            off(eng.flat), off(eng.gflat), off(self.m), off(self.v), off(self.ema) if self.ema is not None else None, hi - lo,
            self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, max(self.step_count, 1),
            self.ema_decay if self.ema_decay is not None else 0.0, grad_scale,
            self.step_dev.data_ptr() if self.step_dev is not None else None, ops._stream()), "adam_ema")
        ops.LAUNCHES += 1

    def ema_state_dict(self, prefix=""):
        """EMA weights under the model's own parameter names."""
        eng = self._ensure()
        out = {}
        for k, (o, n) in eng.offs.items():
            out[prefix + k] = self.ema[o:o + n].view(eng.named[k].shape).clone()
        return out

    # ------------------------------------------------------------------ checkpoint interop
    def _order(self):
        """Parameter names in ``model.parameters()`` order — the index space of torch.optim.Adam's state_dict."""
        return [k for k, _ in self.unet.named_parameters()]

    def state_dict(self):
        """The dict ``torch.optim.Adam(model.parameters(), ...).state_dict()`` would hold at this point
        (trainers/ddpm_trainer.py:139-143,873), so a checkpoint written here resumes in the reference trainer and
        vice versa: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` cut out of the flat moment arenas."""
        names = self._order()
        state = {}
        if self.step_count > 0:
            eng = self._ensure()
            for i, k in enumerate(names):
                o, n = eng.offs[k]
                shape = eng.named[k].shape
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.m[o:o + n].view(shape).clone(),
                            "exp_avg_sq": self.v[o:o + n].view(shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(names)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd, device=None):
        """Inverse of ``state_dict``: accepts the reference trainer's ``optimizer_state_dict`` (trainers/ddpm_trainer.py:917)."""
        groups = sd["param_groups"]
        names = self._order()
        if len(groups) != 1 or list(groups[0]["params"]) != list(range(len(names))):
            raise ValueError(f"expected one parameter group over {len(names)} parameters (the reference's Adam), got "
                             f"{[len(g['params']) for g in groups]}")
        g = groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("decoupled_weight_decay"):
            raise ValueError("amsgrad / maximize / decoupled weight decay are not implemented by the fused Adam")
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), tuple(g["betas"]), float(g["eps"]), float(g["weight_decay"])
        state = sd["state"]
        self.step_dev = None          # rebuilt from step_count by the next begin_step()
        if not state:
            self.step_count, self.m, self.v = 0, None, None
            return
        if sorted(state) != list(range(len(names))):
            raise ValueError("optimizer state must cover every parameter")
        steps = {int(float(st["step"])) for st in state.values()}
        if len(steps) != 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): one fused launch applies one bias correction")
        eng = self.unet.engine
        if eng.flat is None:
            if device is None:
                device = next(self.unet.parameters()).device
            eng.prepare(device)
        ema = self.ema
        self.m = None
        self._ensure()
        if ema is not None and ema.numel() == eng.flat.numel() and ema.device == eng.flat.device:
            self.ema = ema                      # keep EMA weights loaded before the optimizer state
        for i, k in enumerate(names):
            o, n = eng.offs[k]
            for key, arena in (("exp_avg", self.m), ("exp_avg_sq", self.v)):
                src = state[i][key]
                if src.numel() != n:
                    raise ValueError(f"{key} of parameter {i} ({k}) has {src.numel()} elements, expected {n}")
                arena[o:o + n].copy_(src.reshape(-1))
        self.step_count = steps.pop()

    def load_ema_state_dict(self, sd, prefix=""):
        """EMA weights from the reference's ``ema_model.state_dict()`` (trainers/ddpm_trainer.py:872,914-915); buffers and
        other keys outside ``prefix`` + parameter name are ignored."""
        if self.ema_decay is None:
            return
        eng = self._ensure()
        for k, (o, n) in eng.offs.items():
            if prefix + k not in sd:
                raise KeyError(f"EMA state has no entry {prefix + k}")
            self.ema[o:o + n].copy_(sd[prefix + k].reshape(-1))
