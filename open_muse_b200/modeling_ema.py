"""``EMAModel`` -- exponential moving average of the trained weights with the reference's surface (muse/modeling_ema.py:8-240:
``step / copy_to / store / restore / to / state_dict / load_state_dict / save_pretrained / from_pretrained``, same decay
schedule).  SURVEY section 8(f)4 "step-adjacent host op": the per-parameter Python loop of the reference
(``s.sub_(one_minus_decay * (s - p))``: three kernels per parameter, ~450 launches for the base model) becomes three
multi-tensor passes over all parameters (``torch._foreach_*``) with bit-identical arithmetic."""
from __future__ import annotations

import copy
from typing import Any, Dict, Iterable, Optional, Union

import torch


class EMAModel:
    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float = 0.9999, min_decay: float = 0.0,
                 update_after_step: int = 0, update_every: int = 1, use_ema_warmup: bool = False,
                 inv_gamma: Union[float, int] = 1.0, power: Union[float, int] = 2 / 3, model_cls: Optional[Any] = None,
                 model_config: Dict[str, Any] = None):
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.temp_stored_params = None
        self.decay, self.min_decay = decay, min_decay
        self.update_after_step, self.update_every = update_after_step, update_every
        self.use_ema_warmup, self.inv_gamma, self.power = use_ema_warmup, inv_gamma, power
        self.optimization_step = 0
        self.cur_decay_value = None
        self.model_cls, self.model_config = model_cls, model_config

    @classmethod
    def from_pretrained(cls, path, model_cls) -> "EMAModel":
        config = model_cls.load_config(path)
        model = model_cls.from_pretrained(path)
        ema = cls(model.parameters(), model_cls=model_cls, model_config=model.config)
        ema.load_state_dict({k: v for k, v in config.items() if k in ema.state_dict() and k != "shadow_params"})
        return ema

    def save_pretrained(self, path):
        if self.model_cls is None:
            raise ValueError("`save_pretrained` can only be used if `model_cls` was defined at __init__.")
        if self.model_config is None:
            raise ValueError("`save_pretrained` can only be used if `model_config` was defined at __init__.")
        cfg = {k: v for k, v in dict(self.model_config).items() if not k.startswith("_")}
        model = self.model_cls(**cfg)
        state = self.state_dict()
        state.pop("shadow_params", None)
        model.register_to_config(**state)
        self.copy_to(model.parameters())
        model.save_pretrained(path)

    def get_decay(self, optimization_step: int) -> float:
        """(1 + step) / (10 + step), or the warm-up power law, clipped to [min_decay, decay] (reference :89-106)."""
        step = max(0, optimization_step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        if self.use_ema_warmup:
            value = 1 - (1 + step / self.inv_gamma) ** -self.power
        else:
            value = (1 + step) / (10 + step)
        return max(min(value, self.decay), self.min_decay)

    @torch.no_grad()
    def step(self, parameters: Iterable[torch.nn.Parameter]):
        parameters = list(parameters)
        self.optimization_step += 1
        if (self.optimization_step - 1) % self.update_every != 0:
            return
        decay = self.get_decay(self.optimization_step)
        self.cur_decay_value = decay
        one_minus_decay = 1 - decay
        train_s = [s for s, p in zip(self.shadow_params, parameters) if p.requires_grad]
        train_p = [p.detach() for p in parameters if p.requires_grad]
        if train_s:  # s -= (1 - decay) * (s - p): the reference's expression, three multi-tensor passes
            delta = torch._foreach_sub(train_s, train_p)
            torch._foreach_mul_(delta, one_minus_decay)
            torch._foreach_sub_(train_s, delta)
        for s, p in zip(self.shadow_params, parameters):
            if not p.requires_grad:
                s.copy_(p)

    @torch.no_grad()
    def copy_to(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        # p.copy_ (not p.data.copy_): the in-place write must bump the parameter's version counter, which is what the
        # models' packed bf16 operand caches key on -- validation with EMA weights (training/train_muse.py:857-908:
        # store / copy_to / validate / restore) would otherwise run every GEMM with the stale non-EMA operands.
        for s, p in zip(self.shadow_params, list(parameters)):
            p.copy_(s.to(p.device))

    def to(self, device=None, dtype=None) -> None:
        self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                              for p in self.shadow_params]

    def state_dict(self) -> dict:
        return {"decay": self.decay, "min_decay": self.min_decay, "optimization_step": self.optimization_step,
                "update_after_step": self.update_after_step, "use_ema_warmup": self.use_ema_warmup,
                "inv_gamma": self.inv_gamma, "power": self.power, "shadow_params": self.shadow_params}

    def store(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        self.temp_stored_params = [p.detach().cpu().clone() for p in parameters]

    def restore(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        if self.temp_stored_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        with torch.no_grad():
            for c, p in zip(self.temp_stored_params, parameters):
                p.copy_(c)  # bumps the version counter (see copy_to)
        self.temp_stored_params = None

    def load_state_dict(self, state_dict: dict) -> None:
        state_dict = copy.deepcopy(state_dict)
        self.decay = state_dict.get("decay", self.decay)
        if self.decay < 0.0 or self.decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        checks = (("min_decay", float, "Invalid min_decay"), ("optimization_step", int, "Invalid optimization_step"),
                  ("update_after_step", int, "Invalid update_after_step"), ("use_ema_warmup", bool, "Invalid use_ema_warmup"),
                  ("inv_gamma", (float, int), "Invalid inv_gamma"), ("power", (float, int), "Invalid power"))
        for name, types, msg in checks:
            value = state_dict.get(name, getattr(self, name))
            if not isinstance(value, types):
                raise ValueError(msg)
            setattr(self, name, value)
        shadow = state_dict.get("shadow_params", None)
        if shadow is not None:
            self.shadow_params = shadow
            if not isinstance(self.shadow_params, list):
                raise ValueError("shadow_params must be a list")
            if not all(isinstance(p, torch.Tensor) for p in self.shadow_params):
                raise ValueError("shadow_params must all be Tensors")
