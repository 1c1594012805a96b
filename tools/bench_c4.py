#!/usr/bin/env python
"""BASELINE config 4: text-conditioned MaskGitTransformer with the cc12m dims (L=22, H=1024, nh=16, I=4096, vocab 8256,
output 8192, cross-attention to 77x768, RMSNorm, no normformer), per-GPU batch 64, bf16, AdamW, DDP when launched under
torchrun.  CUDA-event timing, max over ranks; rank 0 prints one JSON line.

    python tools/bench_c4.py [--steps 5 --warmup 3 --batch 64]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_c4.py
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_muse_b200 import MaskGitTransformer, ops  # noqa: E402

CFG = dict(vocab_size=8256, hidden_size=1024, intermediate_size=4096, num_hidden_layers=22, num_attention_heads=16,
           max_position_embeddings=256, encoder_hidden_size=768, add_cross_attention=True,
           project_encoder_hidden_states=False, codebook_size=8192, num_vq_tokens=256, norm_type="rmsnorm",
           layer_norm_eps=1e-6, use_normformer=False, use_encoder_layernorm=True, use_bias=False, hidden_dropout=0.0,
           attention_dropout=0.0, use_codebook_size_for_output=True)
TRAIN_GFLOP_PER_SAMPLE = 691.3  # SURVEY 8d (230.44 fwd x 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--model", default="v1", choices=["v1", "uvit", "uvit512"],
                    help="v1: MaskGitTransformer with the cc12m dims (BASELINE config 4); uvit: MaskGiTUViT_v2 at the reference "
                         "defaults (what configs/cc12m_uvit_clip.yaml's architecture: 'uvit' instantiates); uvit512: the same "
                         "with force_down_up_sample=True on 32x32 = 1024 tokens (configs/research_run_512_with_downsample.yaml)")
    args = ap.parse_args()
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    uvit = args.model in ("uvit", "uvit512")
    n_tok = 1024 if args.model == "uvit512" else 256
    if uvit:
        from open_muse_b200 import MaskGiTUViT_v2

        model = MaskGiTUViT_v2(force_down_up_sample=args.model == "uvit512").to(dev).train()
        with torch.no_grad():  # leave the zero-init regime so that every branch carries signal
            for k, x in model.state_dict().items():
                if "adaLN_modulation.mapper" in k or k.endswith("gamma") or k.endswith("beta") or k == "mlm_layer.conv1.weight":
                    x.normal_(0, 0.02)
    else:
        model = MaskGitTransformer(**CFG).to(dev).train()
    n_params = sum(p.numel() for p in model.parameters())
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    B, mask_id = args.batch, CFG["vocab_size"] - 1
    enc = torch.randn(B, 77, 768, device=dev, generator=g)
    pooled = torch.randn(B, 768, device=dev, generator=g)
    micro = torch.tensor([[256.0, 256.0, 0.0, 0.0, 6.0]], device=dev).repeat(B, 1)

    def step(i):
        ids = torch.randint(0, 8192, (B, n_tok), device=dev, generator=g)
        t = torch.rand(B, device=dev, generator=g)
        n_mask = (n_tok * torch.cos(t * torch.pi * 0.5)).round().clamp(min=1)  # train_muse.py:149-226 recipe
        perm = torch.rand(B, n_tok, device=dev, generator=g).argsort(dim=-1)
        mask = perm < n_mask[:, None]
        inp = torch.where(mask, mask_id, ids)
        lab = torch.where(mask, ids, -100)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if uvit:
                _, loss = net(inp, enc, pooled, micro, labels=lab)
            else:
                _, loss = net(inp, encoder_hidden_states=enc, labels=lab)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / args.steps
    if rank == 0:
        print(json.dumps({
            "metric": "samples/sec text2image train step: " + (
                {"uvit": "MaskGiTUViT_v2 (reference defaults, 256 tokens)",
                 "uvit512": "MaskGiTUViT_v2 (force_down_up_sample, 1024 tokens outside / 256 inside)"}[args.model]
                if uvit else "MaskGitTransformer (cc12m dims, config 4)"),
            "value": B * world / (ms * 1e-3), "unit": "samples/s", "n_gpus": world, "ms_per_step": ms, "steps": args.steps,
            "per_gpu_batch": B, "params_m": n_params / 1e6, "loss": float(loss),
            "tflops_per_gpu_model": (TRAIN_GFLOP_PER_SAMPLE if args.model == "v1" else 3 * 266.0) * B / ms if args.model != "uvit512" else None, "gpu_launches_per_step": (ops.launches() - l0) // args.steps,
            "grad_allreduce_gb": n_params * 4 / 1e9, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
