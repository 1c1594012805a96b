#!/usr/bin/env python
"""One pass of a secondary hot path between cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists.
  PROFILE_PATH=decode   generate2, base-256 model, B=64, 12 steps, kernels launched one by one (config 5)
  PROFILE_PATH=encode   MaskGitVQGAN f16-256 get_code, B=64, conv precision from PROFILE_CONV (bf16x3 | bf16)
  PROFILE_PATH=vqdecode MaskGitVQGAN decode_code_uint8, B=64
  PROFILE_PATH=uvit     one MaskGiTUViT_v2 (class defaults) decode-step forward, 16 x 256 tokens
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

path = os.environ.get("PROFILE_PATH", "decode")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if path == "decode":
    from open_muse_b200.modeling_transformer import MaskGitTransformer

    model = MaskGitTransformer(**bench.BASE_CFG).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(7)
    cls = torch.randint(0, 1000, (64,), device=dev)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            model.generate2(class_ids=cls.clone(), timesteps=12, generator=gen, use_cuda_graph=False)
elif path in ("encode", "vqdecode"):
    from open_muse_b200 import MaskGitVQGAN

    vq = MaskGitVQGAN().to(dev).eval()
    vq.set_conv_precision(os.environ.get("PROFILE_CONV", "bf16x3"))
    pix = torch.rand(64, 3, 256, 256, device=dev)
    ids = torch.randint(0, 1024, (64, 256), device=dev)
    run = (lambda: vq.get_code(pix)) if path == "encode" else (lambda: vq.decode_code_uint8(ids))
else:
    from open_muse_b200 import MaskGiTUViT_v2

    m = MaskGiTUViT_v2().to(dev).eval()
    ids = torch.randint(0, 8192, (16, 256), device=dev)
    enc, ce = torch.randn(16, 77, 768, device=dev), torch.randn(16, 768, device=dev)
    mc = torch.tensor([[512.0, 512.0, 0.0, 0.0, 6.0]], device=dev).repeat(16, 1)

    def run():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            m(ids, enc, ce, mc)

for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", path)
