#!/usr/bin/env python
"""One MaskGitVQGAN f16-256 encode + decode_code at batch VQ_BATCH inside a cudaProfiler range (for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum ... python tools/profile_vqgan.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_muse_b200 import MaskGitVQGAN  # noqa: E402

torch.manual_seed(0)
v = MaskGitVQGAN().to("cuda").eval()
B = int(os.environ.get("VQ_BATCH", "32"))
img = torch.rand(B, 3, 256, 256, device="cuda")
ids = v.get_code(img)
v.decode_code(ids)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ids = v.get_code(img)
torch.cuda.synchronize()
rec = v.decode_code(ids)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
