#!/usr/bin/env python
"""One base-256 train step (B=256) between cudaProfilerStart/Stop, for `ncu --profile-from-start off`.
Usage under gpurun:
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from open_muse_b200.modeling_transformer import MaskGitTransformer  # noqa: E402

B = int(os.environ.get("PROFILE_BATCH", "256"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = MaskGitTransformer(**bench.BASE_CFG).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
gen = torch.Generator(device=dev).manual_seed(1)
tok = torch.randint(0, 1024, (B, 256), device=dev)
cls = torch.randint(0, 1000, (B,), device=dev)


def step():
    inp, lab = bench.mask_batch(tok, cls, 2024, 1024, gen=gen)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, loss = model(inp, labels=lab)
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(int(os.environ.get("PROFILE_WARMUP", "3"))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(int(os.environ.get("PROFILE_STEPS", "1"))):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step done")
