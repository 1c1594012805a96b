#!/usr/bin/env python
"""bench.py -- masked-token train step (MaskGitTransformer base-256) on N B200s, one JSON line on rank 0.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference algorithm on the host CPU (oracle port, fp32 eager)

Workload = BASELINE.json configs[1]: MaskGitTransformer base (8x512, nh 8, I 2048, seq 257, vocab 2025),
class-conditional, bf16 compute, per-GPU batch 256 of synthetic pre-tokenised ImageNet-shaped samples
(random VQ ids + random class ids; the reference supports pre-encoded tokens and notes that images
"may be pre-encoded for faster training", README.md:137).  A step = the reference loop body
(training/train_maskgit_imagenet.py:403-452): masking recipe -> forward (logits, loss) -> backward ->
AdamW step -> zero_grad.  `value` times it with inputs resident in HBM; `e2e` adds the per-step
host->device copy of the token batch from pinned memory and a device->host read of the loss.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BASE_CFG = dict(vocab_size=2025, max_position_embeddings=257, hidden_size=512, num_hidden_layers=8,
                num_attention_heads=8, intermediate_size=2048, codebook_size=1024, num_vq_tokens=256, num_classes=1000,
                hidden_dropout=0.0, attention_dropout=0.0)
PER_GPU_BATCH = 256
FWD_GFLOP_PER_IMG = 18.997  # SURVEY 8d / BASELINE.md section 3 (17.915 linear + 1.082 attention core)
METRIC = "images/sec masked-token train step (base-256)"


def mask_batch(tokens, class_ids, mask_id, codebook, gen=None):
    """prepare_inputs_and_labels without the VQ encode (training/train_maskgit_imagenet.py:375-393)."""
    B, S = tokens.shape
    dev = tokens.device
    timesteps = torch.rand(B, device=dev, generator=gen)
    mask_prob = torch.cos(timesteps * math.pi * 0.5).clip(0.0)
    n_mask = (S * mask_prob).round().clamp(min=1)
    perm = torch.rand(B, S, device=dev, generator=gen).argsort(dim=-1)
    mask = perm < n_mask.unsqueeze(-1)
    input_ids = torch.where(mask, mask_id, tokens)
    labels = torch.where(mask, tokens, -100)
    input_ids = torch.cat([(class_ids + codebook).unsqueeze(-1), input_ids], dim=-1)
    labels = torch.cat([torch.full((B, 1), -100, device=dev, dtype=labels.dtype), labels], dim=-1)
    return input_ids, labels


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def gemm_traffic_from_profile():
    """Average DRAM bytes per tcgen05-GEMM launch from the committed `ncu --set full` capture (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_ncu_full_gemm.txt")
    if not os.path.exists(path):
        return None
    tot, n = 0.0, 0
    for line in open(path):
        if "gemm_tcgen05_kernel" not in line:
            continue
        try:
            rd = float(line.split("dram_read=")[1].split("MB")[0])
            wr = float(line.split("dram_write=")[1].split("MB")[0])
            tot += (rd + wr) * 1e6
            n += 1
        except Exception:
            pass
    return tot / n if n else None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("hbm_gbs"), "measured (MEASURED_PEAKS.json, sustained bf16)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md: ~1.4 PF sustained, 6.65 TB/s)"


def cpu_reference_step_fn(batch, device="cpu"):
    """The reference algorithm (oracle port, torch eager): fwd + bwd + AdamW on `batch` samples.  fp32 on the host CPU
    (the reference arm); with device="cuda" the same eager op sequence under bf16 autocast on the GPU -- what the
    reference's own PyTorch modules would run on this B200 (`--impl reference --ref-device cuda`, informational)."""
    from oracle import transformer_oracle as T
    from open_muse_b200.modeling_transformer import MaskGitTransformer

    torch.manual_seed(0)
    params = {k: v.clone().to(device).requires_grad_(True) for k, v in MaskGitTransformer(**BASE_CFG).state_dict().items()}
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=0.01, fused=(device != "cpu"))
    g = torch.Generator(device=device).manual_seed(1)

    def step():
        tokens = torch.randint(0, 1024, (batch, 256), generator=g, device=device)
        cls = torch.randint(0, 1000, (batch,), generator=g, device=device)
        inp, lab = mask_batch(tokens, cls, 2024, 1024, gen=g)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(device != "cpu")):
            _, loss = T.forward(params, BASE_CFG, inp, labels=lab)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return float(loss)

    return step


CPU_SAMPLE_BATCH = 32


def cpu_threads():
    """Threads for the CPU arm: every host core up to 64 (torch-eager fp32 at batch 32 stops scaling, and on a
    128-core box oversubscribing all of them was 10x SLOWER than 8 threads in the first measured run)."""
    return max(1, min(os.cpu_count() or 1, 64))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    torch.set_num_threads(cores)
    on_gpu = args.ref_device == "cuda"
    batch = args.batch if on_gpu else CPU_SAMPLE_BATCH
    step = cpu_reference_step_fn(batch, args.ref_device)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    if on_gpu:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()  # returns float(loss): synchronises every step
    dt = (time.perf_counter() - t0) / args.steps
    value = batch / dt
    sample = f"{args.steps} steps x batch {batch} of the base-256 train step (fwd+bwd+AdamW), " + (
        "bf16-autocast torch eager on cuda:0 (informational)" if on_gpu else "fp32 torch eager")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if on_gpu else "f32", "data": "synthetic",
        "ref_device": args.ref_device,
        "config": {"workload": f"MaskGitTransformer base (8x512, seq 257, vocab 2025) class-cond train step, {args.ref_device} sample batch {batch}",
                   "global_batch": batch, "seq_len": 257},
        "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH, help="per-GPU batch (default = BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-step", action="store_true", help="skip the extra 'train step incl. VQ encode' measurement")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="launch the kernels of each step one by one instead of replaying the captured step (N=1 default: graph)")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference only: cpu (the reference arm) or cuda (same eager ops under bf16 autocast, informational)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from open_muse_b200 import ops
    from open_muse_b200.modeling_transformer import MaskGitTransformer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    B = args.batch

    torch.manual_seed(0)
    model = MaskGitTransformer(**BASE_CFG).to(dev).train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
    use_graph = world == 1 and not args.no_cuda_graph
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True, capturable=use_graph)
    gen = None if use_graph else torch.Generator(device=dev).manual_seed(100 + rank)  # graph capture: default generator
    torch.cuda.manual_seed(100 + rank)
    n_buf = 4
    host_tok = [torch.randint(0, 1024, (B, 256), generator=torch.Generator().manual_seed(1000 * rank + i)).pin_memory() for i in range(n_buf)]
    host_cls = [torch.randint(0, 1000, (B,), generator=torch.Generator().manual_seed(2000 * rank + i)).pin_memory() for i in range(n_buf)]
    dev_tok = [t.to(dev) for t in host_tok]
    dev_cls = [t.to(dev) for t in host_cls]

    def step(tokens, cls):
        inp, lab = mask_batch(tokens, cls, 2024, 1024, gen=gen)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, loss = net(inp, labels=lab)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n

    for i in range(warmup):
        step(dev_tok[i % n_buf], dev_cls[i % n_buf])
    torch.cuda.synchronize()
    l_eager = ops.launches()
    step(dev_tok[0], dev_cls[0])
    launches_per_step = ops.launches() - l_eager  # kernels of this package per step (a replayed graph launches the same)
    run = step
    if use_graph:  # the whole step (masking, fwd, loss, bwd, AdamW) captured once, replayed per step
        from open_muse_b200.graphs import GraphedStep

        try:
            run = GraphedStep(step, (dev_tok[0], dev_cls[0]), warmup=2)
            for i in range(2):
                run(dev_tok[i % n_buf], dev_cls[i % n_buf])
        except Exception as e:  # same kernels launched one by one (reported in config.cuda_graph)
            print(f"bench: CUDA graph capture failed ({type(e).__name__}: {e}); timing the eager launch path", file=sys.stderr)
            torch.cuda.synchronize()
            run, use_graph = step, False
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev = timed(lambda i: run(dev_tok[i % n_buf], dev_cls[i % n_buf]), args.steps)
    launches = launches_per_step * args.steps

    def e2e_step(i):
        tok = host_tok[i % n_buf].to(dev, non_blocking=True)
        cls = host_cls[i % n_buf].to(dev, non_blocking=True)
        return float(run(tok, cls))  # .item(): device->host read of the loss

    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    model._packed.key = None  # graph replays update the parameters without bumping their version counters: re-pack for eager use

    # ---- roofline of the dominant kernel (tcgen05 GEMM), measured live with CUDA events around every launch
    peak_tf, peak_hbm, peak_src = measured_peaks()
    ops.profile_gemms(True)
    for i in range(2):
        step(dev_tok[i % n_buf], dev_cls[i % n_buf])
    torch.cuda.synchronize()
    gemm_ms, gemm_flops, gemm_n = ops.profile_gemms(False)
    roof = None
    if gemm_n:
        ach = gemm_flops / (gemm_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all Linear fwd/dgrad/wgrad GEMMs of the step)",
                "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                "traffic": gemm_traffic_from_profile(), "algorithmic_flops_per_launch": gemm_flops / gemm_n,
                "peak_source": peak_src, "launches_per_step": gemm_n // 2,
                "gemm_share_of_step": (gemm_ms / 2) / ms_dev}

    # ---- BASELINE config 2 (ii): the same train step fed from pixels, i.e. including the frozen MaskGitVQGAN tokeniser
    # (train_maskgit_imagenet.py:357-369: fp32, no autocast) in front of the transformer step.  N=1 only, informational.
    full = None
    if world == 1 and not args.no_full_step:
        from open_muse_b200 import MaskGitVQGAN

        torch.manual_seed(1)
        vq = MaskGitVQGAN().to(dev).eval()
        pix = torch.rand(B, 3, 256, 256, device=dev)
        chunk = 64

        def tokenise():
            return torch.cat([vq.get_code(pix[j:j + chunk]) for j in range(0, B, chunk)])

        def full_step(i):
            return step(tokenise(), dev_cls[i % n_buf])

        full_step(0)
        ms_full = timed(full_step, 3)
        ms_enc = timed(lambda i: tokenise(), 3)
        full = {"value": B / (ms_full * 1e-3), "unit": "images/s", "ms_per_step": ms_full, "vq_encode_ms": ms_enc,
                "vq_encode_images_per_s": B / (ms_enc * 1e-3),
                "note": "pixels resident in HBM -> MaskGitVQGAN f16-256 get_code (fp32-faithful bf16x3 tcgen05 convs) -> "
                        "masking -> train step; random-init tokeniser"}
        del vq, pix
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads()
        torch.set_num_threads(cores)
        cstep = cpu_reference_step_fn(CPU_SAMPLE_BATCH)
        cstep()
        t0 = time.perf_counter()
        n = 0
        while n < 2 or (time.perf_counter() - t0 < 15 and n < 12):
            cstep(); n += 1
        cdt = (time.perf_counter() - t0) / n
        cpu = {"value": CPU_SAMPLE_BATCH / cdt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{n} steps x batch {CPU_SAMPLE_BATCH} of the same train step, oracle fp32 torch eager on {cores} host threads"}

    if rank == 0:
        gb = B * world
        out = {
            "metric": METRIC, "value": gb / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "MaskGitTransformer base (8x512, nh8, I2048, seq 257, vocab 2025) class-cond train step: "
                                   "masking + fwd + loss + bwd + AdamW, pre-tokenised synthetic ImageNet batch",
                       "global_batch": gb, "per_gpu_batch": B, "seq_len": 257,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "l2": "no explicit flush: per-step working set (~15 GB activations + 0.5 GB weights/grads/optimizer) >> 126 MB L2",
                       "gemm_backend": os.environ.get("MUSE_B200_GEMM", "tcgen05"),
                       "cuda_graph": bool(use_graph)},
            "e2e": {"value": gb / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": world * (B * 256 * 8 + B * 8), "d2h_bytes_per_step": world * 4},
            "gpu_launches": launches,
            "tflops_per_gpu_model": 3 * FWD_GFLOP_PER_IMG * B / ms_dev,
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "full_step_incl_vq_encode": full,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
