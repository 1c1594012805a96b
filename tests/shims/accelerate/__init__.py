"""Minimal stand-in for huggingface/accelerate (test infrastructure; see tests/shims/README.md)."""
import contextlib
import os

import torch

from . import logging, utils  # noqa: F401
from .utils import DistributedType


class _State:
    def __init__(self, acc):
        self.acc = acc
        self.deepspeed_plugin = None

    def __repr__(self):
        a = self.acc
        return (f"Distributed environment: {a.distributed_type}\nNum processes: {a.num_processes}\n"
                f"Process index: {a.process_index}\nDevice: {a.device}\nMixed precision type: {a.mixed_precision}")


class _AutocastForward(torch.nn.Module):
    """accelerate's prepare(): forward runs under autocast and its floating-point outputs are converted to fp32."""

    def __init__(self, module, dtype):
        super().__init__()
        self.module = module
        self._dtype = dtype

    def forward(self, *args, **kwargs):
        with torch.autocast("cuda", dtype=self._dtype):
            out = self.module(*args, **kwargs)

        def up(t):
            return t.float() if torch.is_tensor(t) and t.is_floating_point() else t

        if isinstance(out, tuple):
            return tuple(up(t) for t in out)
        return up(out)

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.module, name)


class _AcceleratedOptimizer:
    """accelerate's optimizer wrapper: ``step`` / ``zero_grad`` only act on the micro-step that closes an accumulation
    window (``accelerator.sync_gradients``); everything else is the wrapped optimizer."""

    def __init__(self, optimizer, accelerator):
        self.__dict__["optimizer"], self.__dict__["_acc"] = optimizer, accelerator

    def step(self, *a, **k):
        if self._acc.sync_gradients:
            return self.optimizer.step(*a, **k)

    def zero_grad(self, *a, **k):
        if self._acc.sync_gradients:
            return self.optimizer.zero_grad(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["optimizer"], name)

    def __setattr__(self, name, value):
        setattr(self.__dict__["optimizer"], name, value)


class _AcceleratedScheduler:
    def __init__(self, scheduler, accelerator):
        self.__dict__["scheduler"], self.__dict__["_acc"] = scheduler, accelerator

    def step(self, *a, **k):
        if self._acc.sync_gradients:
            return self.scheduler.step(*a, **k)

    def __getattr__(self, name):
        return getattr(self.__dict__["scheduler"], name)


class Accelerator:
    last = None  # the most recently constructed instance (tests inspect what the script logged)

    def __init__(self, gradient_accumulation_steps=1, mixed_precision=None, log_with=None, logging_dir=None,
                 project_dir=None, split_batches=False, **kwargs):
        self.gradient_accumulation_steps = int(gradient_accumulation_steps)
        self.mixed_precision = mixed_precision or "no"
        self.num_processes = int(os.environ.get("WORLD_SIZE", "1"))
        self.process_index = int(os.environ.get("RANK", "0"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        self.distributed_type = DistributedType.MULTI_GPU if self.num_processes > 1 else DistributedType.NO
        use_cuda = torch.cuda.is_available() and os.environ.get("ACCELERATE_USE_CPU", "0") != "1"
        self.device = torch.device("cuda", self.local_process_index) if use_cuda else torch.device("cpu")
        if use_cuda:
            torch.cuda.set_device(self.device)
        if self.num_processes > 1 and not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl" if use_cuda else "gloo")
        self.state = _State(self)
        self.sync_gradients = True
        self.scaler = None
        self._step = 0
        self._models, self._optimizers, self._schedulers = [], [], []
        self._save_hooks, self._load_hooks = [], []
        self.logged = []  # (values, step) pairs seen by log(): tests read them
        self.trackers_initialised = None
        Accelerator.last = self

    # ---- process topology
    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def print(self, *a, **k):
        if self.is_local_main_process:
            print(*a, **k)

    def wait_for_everyone(self):
        if self.num_processes > 1:
            torch.distributed.barrier()

    # ---- prepare / unwrap
    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if self.num_processes > 1:
                    o = torch.nn.parallel.DistributedDataParallel(
                        o, device_ids=[self.device.index] if self.device.type == "cuda" else None)
                if self.mixed_precision in ("bf16", "fp16") and self.device.type == "cuda":
                    o = _AutocastForward(o, torch.bfloat16 if self.mixed_precision == "bf16" else torch.float16)
                self._models.append(o)
            elif isinstance(o, torch.optim.Optimizer):
                self._optimizers.append(o)
                o = _AcceleratedOptimizer(o, self)
            elif hasattr(o, "get_last_lr"):
                self._schedulers.append(o)
                o = _AcceleratedScheduler(o, self)
            out.append(o)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model):
        while isinstance(model, (_AutocastForward, torch.nn.parallel.DistributedDataParallel)):
            model = model.module
        return model

    # ---- the step
    @contextlib.contextmanager
    def accumulate(self, model):
        self._step += 1
        self.sync_gradients = self._step % self.gradient_accumulation_steps == 0
        inner = model.module if isinstance(model, _AutocastForward) else model
        if not self.sync_gradients and isinstance(inner, torch.nn.parallel.DistributedDataParallel):
            with inner.no_sync():
                yield
        else:
            yield

    def backward(self, loss, **kwargs):
        (loss / self.gradient_accumulation_steps).backward(**kwargs)

    def gather(self, tensor):
        if self.num_processes == 1:
            return tensor
        out = [torch.empty_like(tensor) for _ in range(self.num_processes)]
        torch.distributed.all_gather(out, tensor.contiguous())
        return torch.cat(out)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        return torch.nn.utils.clip_grad_norm_(list(parameters), max_norm, norm_type=norm_type)

    # ---- tracking
    def init_trackers(self, project_name, config=None, init_kwargs=None):
        self.trackers_initialised = dict(project=project_name, config=config, init_kwargs=init_kwargs)

    def log(self, values, step=None):
        self.logged.append((dict(values), step))

    def end_training(self):
        pass

    # ---- checkpoints
    def get_state_dict(self, model):
        return self.unwrap_model(model).state_dict()

    def save(self, obj, f):
        if self.is_main_process:
            torch.save(obj, f)

    def register_save_state_pre_hook(self, hook):
        self._save_hooks.append(hook)

    def register_load_state_pre_hook(self, hook):
        self._load_hooks.append(hook)

    def save_state(self, output_dir):
        output_dir = str(output_dir)
        if self.is_main_process:
            os.makedirs(output_dir, exist_ok=True)
            weights = [self.unwrap_model(m).state_dict() for m in self._models]
            for h in self._save_hooks:
                h([self.unwrap_model(m) for m in self._models], weights, output_dir)
            for i, w in enumerate(weights):
                torch.save(w, os.path.join(output_dir, f"pytorch_model{'' if i == 0 else '_' + str(i)}.bin"))
            for i, o in enumerate(self._optimizers):
                torch.save(o.state_dict(), os.path.join(output_dir, f"optimizer{'' if i == 0 else '_' + str(i)}.bin"))
            for i, s in enumerate(self._schedulers):
                torch.save(s.state_dict(), os.path.join(output_dir, f"scheduler{'' if i == 0 else '_' + str(i)}.bin"))
        self.wait_for_everyone()
        return output_dir

    def load_state(self, input_dir):
        input_dir = str(input_dir)
        for h in self._load_hooks:
            h([self.unwrap_model(m) for m in self._models], input_dir)
        for i, m in enumerate(self._models):
            p = os.path.join(input_dir, f"pytorch_model{'' if i == 0 else '_' + str(i)}.bin")
            self.unwrap_model(m).load_state_dict(torch.load(p, map_location="cpu"))
        for i, o in enumerate(self._optimizers):
            o.load_state_dict(torch.load(os.path.join(input_dir, f"optimizer{'' if i == 0 else '_' + str(i)}.bin"), map_location="cpu"))
        for i, s in enumerate(self._schedulers):
            s.load_state_dict(torch.load(os.path.join(input_dir, f"scheduler{'' if i == 0 else '_' + str(i)}.bin")))
