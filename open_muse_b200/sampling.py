"""Mask schedules and sampling helpers with the reference's public names (muse/sampling.py:9-77).

These are the host-side scalar/torch helpers the training scripts import
(``from muse.sampling import cosine_schedule``; training/train_maskgit_imagenet.py:375-378).  The
per-step sample / confidence / re-mask work of ``generate2`` runs in the fused CUDA kernel
(csrc/sample.cu); the functions below define its semantics and serve user code that calls them directly.
"""
from __future__ import annotations

import math
from functools import partial

import torch


def log(t, eps=1e-20):
    """log with the argument clamped from below (sampling.py:9-10)."""
    return t.clamp(min=eps).log()


def gumbel_noise(t, generator=None):
    """-log(-log(U)), U ~ uniform(0,1) drawn with ``generator`` in the shape/dtype/device of ``t`` (:13-15)."""
    u = torch.zeros_like(t).uniform_(0, 1, generator=generator)
    return -log(-log(u))


def gumbel_sample(t, temperature=1.0, dim=-1, generator=None):
    return (t / max(temperature, 1e-10) + gumbel_noise(t, generator=generator)).argmax(dim=dim)


def top_k(logits, thres=0.9):
    """Keeps the ceil((1-thres)*V) largest logits per position, -inf elsewhere (:22-27)."""
    k = math.ceil((1 - thres) * logits.shape[-1])
    values, index = logits.topk(k, dim=-1)
    return torch.full_like(logits, float("-inf")).scatter_(2, index, values)


def mask_by_random_topk(mask_len, probs, temperature=1.0, generator=None):
    """True where log p + temperature * gumbel is below the row's ``mask_len``-th smallest value (:30-35)."""
    confidence = log(probs) + temperature * gumbel_noise(probs, generator=generator)
    ranked = confidence.sort(dim=-1).values
    threshold = ranked.gather(1, mask_len.long())
    return confidence < threshold


def cosine_schedule(t):
    return torch.cos(t * math.pi * 0.5)


def linear_schedule(t):
    return (1 - t).clamp(min=1e-6, max=1.0)


def pow(t, method):
    exponent = float(method.replace("pow", ""))
    return (1.0 - t**exponent).clamp(min=1e-6, max=1.0)


def sigmoid_schedule(t, start=-3, end=3, tau=1.0, clip_min=1e-6):
    v_start = torch.sigmoid(torch.tensor(start / tau))
    v_end = torch.sigmoid(torch.tensor(end / tau))
    out = torch.sigmoid((t * (end - start) + start) / tau)
    return torch.clip((v_end - out) / (v_end - v_start), clip_min, 1.0)


def get_mask_chedule(method, **schedule_kwargs):  # (sic) name kept: it is the reference's public symbol
    if method == "cosine":
        return cosine_schedule
    if method == "linear":
        return linear_schedule
    if "pow" in method:
        return partial(pow, method=method)
    if method == "sigmoid":
        return partial(sigmoid_schedule, **schedule_kwargs)
    raise ValueError("Unknown schedule method: {}".format(method))
