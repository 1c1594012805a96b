"""CUDA-graph capture of a whole training (or inference) step.

The kernels of this package are launched through a C ABI on torch's current stream, build their TMA descriptors on the
host and never synchronise, so a complete step -- masking, forward, loss, backward, fused AdamW -- can be captured once and
replayed: the ~330 launches of a base-256 train step then cost one graph launch instead of ~1.5-2 ms of launch gaps.

    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, fused=True, capturable=True)   # capturable optimizer state
    def step(tokens, class_ids):
        ...forward under autocast, loss.backward(), opt.step(), opt.zero_grad(set_to_none=True)
        return loss
    graphed = GraphedStep(step, (tokens, class_ids))
    loss = graphed(tokens, class_ids)        # copies the inputs into the static buffers and replays

Rules (the usual ones of ``torch.cuda.graph``): fixed shapes, no host synchronisation inside ``fn``, random numbers from the
default CUDA generator (graph-safe), optimizer created with ``capturable=True``.  DDP steps: construct DDP on a side stream,
set TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group, warm up >= 11 iterations and capture with
``capture_error_mode="thread_local"`` (bench.py --ddp-graph 1 does exactly that; the NCCL bucket all-reduces become graph nodes).
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch


class GraphedStep:
    def __init__(self, fn: Callable, example_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 capture_error_mode: str = "global"):
        self.static_inputs = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # lazy initialisation (kernel attributes, caches, optimizer state) off the capture
            for _ in range(max(1, warmup)):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.static_output = fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_output
