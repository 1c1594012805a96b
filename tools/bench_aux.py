#!/usr/bin/env python
"""Secondary measurements (BASELINE configs 3 and 5): MaskGitVQGAN f16-256 encode->decode at B=128 and
generate2 12-step decode at B=64 on the base model.  CUDA-event timings, printed as JSON lines."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from open_muse_b200 import MaskGitTransformer, MaskGitVQGAN, ops  # noqa: E402

dev = "cuda"


def timed(fn, n, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


which = sys.argv[1:] or ["generate2", "vqgan"]
if "generate2" in which:
    torch.manual_seed(0)
    m = MaskGitTransformer(**bench.BASE_CFG).to(dev).eval()
    gen = torch.Generator(device=dev).manual_seed(7)

    def run():
        cls = torch.randint(0, 1000, (64,), device=dev)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return m.generate2(class_ids=cls, timesteps=12, generator=gen)

    l0 = ops.launches()
    ms = timed(run, 5, warm=2)
    print(json.dumps({"metric": "generate2 decode steps/sec (base model, B=64, 256 tokens, 12 steps)", "value": 12 / (ms * 1e-3),
                      "unit": "steps/s", "ms_per_call": ms, "images_per_s": 64 / (ms * 1e-3),
                      "launches_per_call": (ops.launches() - l0) // 7}), flush=True)
    del m
if "vqgan" in which:
    torch.manual_seed(0)
    v = MaskGitVQGAN().to(dev).eval()
    B = int(os.environ.get("VQ_BATCH", "128"))
    img = torch.rand(B, 3, 256, 256, device=dev)
    ids = v.get_code(img)
    ms_enc = timed(lambda: v.get_code(img), 2, warm=1)
    ms_dec = timed(lambda: v.decode_code(ids), 2, warm=1)
    print(json.dumps({"metric": f"MaskGitVQGAN f16-256 encode / decode_code (fp32, B={B})", "encode_images_per_s": B / (ms_enc * 1e-3),
                      "decode_images_per_s": B / (ms_dec * 1e-3), "encode_ms": ms_enc, "decode_ms": ms_dec,
                      "encode_tflops": 128.76 * B / ms_enc, "decode_tflops": 186.55 * B / ms_dec,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
if "uvit" in which:
    # MaskGiTUViT_v2 at the reference defaults (603 M params: 22 x 1024 + 768-wide U blocks), the model benchmark/muse_perf.py
    # times: 12-step CFG generate2 for 256 px (256 tokens), text states 77 x 768, batch 1 and 8 (doubled by guidance).
    from open_muse_b200 import MaskGiTUViT_v2

    torch.manual_seed(0)
    u = MaskGiTUViT_v2().to(dev).eval()
    n_params = sum(p.numel() for p in u.parameters())
    for bs in (1, 8):
        g = torch.Generator(device=dev).manual_seed(3)
        enc = torch.randn(bs, 77, 768, device=dev)
        pooled = torch.randn(bs, 768, device=dev)
        micro = torch.tensor([[256.0, 256.0, 0.0, 0.0, 6.0]], device=dev)
        e_enc, e_pool = torch.randn(1, 77, 768, device=dev), torch.randn(1, 768, device=dev)

        def run_uvit():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return u.generate2(enc, pooled, micro, e_enc, e_pool, timesteps=12, guidance_scale=8.0, temperature=(2.0, 0.0),
                                   generator=g)

        l0 = ops.launches()
        ms = timed(run_uvit, 3, warm=2)
        print(json.dumps({"metric": f"MaskGiTUViT_v2 generate2 latency (defaults, {n_params / 1e6:.0f} M params, bs {bs}, 12 steps, "
                                    "256 tokens, CFG)", "value": ms, "unit": "ms", "images_per_s": bs / (ms * 1e-3),
                          "ms_per_step": ms / 12, "launches_per_call": (ops.launches() - l0) // 5}), flush=True)
    del u
if "taming" in which:
    # the taming VQGANModel at the geometry of the text-to-image configs (f16, 256 px, 8192 codes, attention at 16x16)
    from open_muse_b200 import VQGANModel

    torch.manual_seed(0)
    t = VQGANModel(num_embeddings=8192).to(dev).eval()
    B = int(os.environ.get("VQ_BATCH", "64"))
    img = torch.rand(B, 3, 256, 256, device=dev)
    ids = t.get_code(img)
    ms_enc = timed(lambda: t.get_code(img), 2, warm=1)
    ms_dec = timed(lambda: t.decode_code(ids), 2, warm=1)
    print(json.dumps({"metric": f"taming VQGANModel f16-256 encode / decode_code (fp32-faithful, B={B})",
                      "encode_images_per_s": B / (ms_enc * 1e-3), "decode_images_per_s": B / (ms_dec * 1e-3),
                      "encode_ms": ms_enc, "decode_ms": ms_dec, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
