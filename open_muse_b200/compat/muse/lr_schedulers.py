"""``muse.lr_schedulers`` for the drop-in package: ``get_scheduler(name, optimizer, num_warmup_steps, num_training_steps)``
as the training scripts call it (training/train_maskgit_imagenet.py:295-300, reference muse/lr_schedulers.py:237-291).
Host-side step-size bookkeeping only (``LambdaLR`` multipliers); the schedule shapes are the standard warm-up families
the reference names: constant, constant_with_warmup, linear, cosine, cosine_with_restarts, polynomial."""
import math
from enum import Enum

from torch.optim.lr_scheduler import LambdaLR

NAMES = ("linear", "cosine", "cosine_with_restarts", "polynomial", "constant", "constant_with_warmup")


class SchedulerType(Enum):
    """reference :29-35"""

    LINEAR = "linear"
    COSINE = "cosine"
    COSINE_WITH_RESTARTS = "cosine_with_restarts"
    POLYNOMIAL = "polynomial"
    CONSTANT = "constant"
    CONSTANT_WITH_WARMUP = "constant_with_warmup"


def _warm(step, warmup):
    return float(step) / float(max(1.0, warmup))


def get_scheduler(name, optimizer, num_warmup_steps=None, num_training_steps=None, num_cycles=1, power=1.0, last_epoch=-1,
                  _lr_end=1e-7, _cosine_cycles=0.5):
    name = getattr(name, "value", name)
    if name not in NAMES:
        raise ValueError(f"{name} is not a valid scheduler; choose one of {NAMES}")
    if name == "constant":
        return LambdaLR(optimizer, lambda _: 1.0, last_epoch=last_epoch)
    if num_warmup_steps is None:
        raise ValueError(f"{name} requires `num_warmup_steps`, please provide that argument.")
    w = num_warmup_steps
    if name == "constant_with_warmup":
        return LambdaLR(optimizer, lambda s: _warm(s, w) if s < w else 1.0, last_epoch=last_epoch)
    if num_training_steps is None:
        raise ValueError(f"{name} requires `num_training_steps`, please provide that argument.")
    n = num_training_steps

    def progress(s):
        return float(s - w) / float(max(1, n - w))

    if name == "linear":
        f = lambda s: _warm(s, w) if s < w else max(0.0, float(n - s) / float(max(1, n - w)))
    elif name == "cosine":
        cycles = _cosine_cycles  # upstream's dispatch does not forward num_cycles to the plain cosine schedule (:283-291)
        f = lambda s: _warm(s, w) if s < w else max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(cycles) * 2.0 * progress(s))))
    elif name == "cosine_with_restarts":
        def f(s):
            if s < w:
                return _warm(s, w)
            p = progress(s)
            return 0.0 if p >= 1.0 else max(0.0, 0.5 * (1.0 + math.cos(math.pi * ((float(num_cycles) * p) % 1.0))))
    else:  # polynomial decay to lr_end = 1e-7
        lr_init, lr_end = optimizer.defaults["lr"], _lr_end
        if not lr_init > lr_end:
            raise ValueError(f"lr_end ({lr_end}) must be be smaller than initial lr ({lr_init})")

        def f(s):
            if s < w:
                return _warm(s, w)
            if s > n:
                return lr_end / lr_init
            return ((lr_init - lr_end) * (1 - (s - w) / (n - w)) ** power + lr_end) / lr_init
    return LambdaLR(optimizer, f, last_epoch)


# the named constructors of the reference module (:38-234), same signatures, on the one implementation above
def get_constant_schedule(optimizer, last_epoch=-1):
    return get_scheduler("constant", optimizer, last_epoch=last_epoch)


def get_constant_schedule_with_warmup(optimizer, num_warmup_steps, last_epoch=-1):
    return get_scheduler("constant_with_warmup", optimizer, num_warmup_steps, last_epoch=last_epoch)


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    return get_scheduler("linear", optimizer, num_warmup_steps, num_training_steps, last_epoch=last_epoch)


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, num_cycles=0.5, last_epoch=-1):
    return get_scheduler("cosine", optimizer, num_warmup_steps, num_training_steps, last_epoch=last_epoch,
                         _cosine_cycles=num_cycles)


def get_cosine_with_hard_restarts_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, num_cycles=1, last_epoch=-1):
    return get_scheduler("cosine_with_restarts", optimizer, num_warmup_steps, num_training_steps, num_cycles=num_cycles,
                         last_epoch=last_epoch)


def get_polynomial_decay_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, lr_end=1e-7, power=1.0, last_epoch=-1):
    return get_scheduler("polynomial", optimizer, num_warmup_steps, num_training_steps, power=power, last_epoch=last_epoch,
                         _lr_end=lr_end)


TYPE_TO_SCHEDULER_FUNCTION = {
    SchedulerType.LINEAR: get_linear_schedule_with_warmup,
    SchedulerType.COSINE: get_cosine_schedule_with_warmup,
    SchedulerType.COSINE_WITH_RESTARTS: get_cosine_with_hard_restarts_schedule_with_warmup,
    SchedulerType.POLYNOMIAL: get_polynomial_decay_schedule_with_warmup,
    SchedulerType.CONSTANT: get_constant_schedule,
    SchedulerType.CONSTANT_WITH_WARMUP: get_constant_schedule_with_warmup,
}
