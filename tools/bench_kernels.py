#!/usr/bin/env python
"""Per-kernel timings at the base-256 shapes (B=256, T=65792, H=512, I=2048) with CUDA events.
Prints achieved GB/s (algorithmic bytes) or TFLOP/s per op -- the inner optimisation loop of the round."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_muse_b200 import ops  # noqa: E402

dev = "cuda"
B, S, H, I, nh = 256, 257, 512, 2048, 8
T = B * S


def timeit(fn, n=10, warm=3):
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


def report(name, ms, gbytes=None, gflop=None):
    s = f"{name:34s} {ms*1e3:9.1f} us"
    if gbytes is not None:
        s += f"  {gbytes/ms:8.1f} GB/s"
    if gflop is not None:
        s += f"  {gflop/ms:8.1f} TFLOP/s"
    print(s, flush=True)


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


only = sys.argv[1:] or None


def want(tag):
    return only is None or any(o in tag for o in only)


u = T * H * 2 / 1e6  # MB of one bf16 [T,H]
if want("glu"):
    ab = bf(T, 2 * I)
    w = torch.ones(I, device=dev)
    dy = bf(T, I)
    dw = torch.zeros(I, device=dev)
    y, st = ops.norm_fwd(ab, w, 1e-6, torch.bfloat16, act=2)
    report("glu+LN fwd (fused)", timeit(lambda: ops.norm_fwd(ab, w, 1e-6, torch.bfloat16, act=2)), 12 * u)
    report("glu+LN bwd (fused)", timeit(lambda: ops.norm_bwd(dy, ab, w, st, torch.bfloat16, dw=dw, act=2)), 20 * u)
    report("glu+LN bwd (fused, saved y)", timeit(lambda: ops.norm_bwd(dy, ab, w, st, torch.bfloat16, dw=dw, act=2, y_fwd=y)), 24 * u)
    report("glu fwd (plain)", timeit(lambda: ops.glu_fwd(ab)), 12 * u)
    report("glu bwd (plain)", timeit(lambda: ops.glu_bwd(ab, dy)), 20 * u)
    gl = ops.glu_fwd(ab)
    y2, st2 = ops.norm_fwd(gl, w, 1e-6, torch.bfloat16)
    report("LN fwd [T,2048] bf16->bf16", timeit(lambda: ops.norm_fwd(gl, w, 1e-6, torch.bfloat16)), 8 * u)
    report("LN bwd [T,2048] bf16", timeit(lambda: ops.norm_bwd(dy, gl, w, st2, torch.bfloat16, dw=dw)), 12 * u)
    del ab, dy, gl, y, y2
if want("norm"):
    x = torch.randn(T, H, device=dev)
    w = torch.ones(H, device=dev)
    xb = bf(T, H)
    dres = torch.randn(T, H, device=dev)
    dw = torch.zeros(H, device=dev)
    y, st = ops.norm_fwd(x, w, 1e-6, torch.bfloat16)
    report("LN fwd fp32->bf16 [T,512]", timeit(lambda: ops.norm_fwd(x, w, 1e-6, torch.bfloat16)), 3 * u)
    report("LN fwd bf16->fp32 +res", timeit(lambda: ops.norm_fwd(xb, w, 1e-6, torch.float32, res=x)), 5 * u)
    report("LN bwd bf16 dy, fp32 x, +dres", timeit(lambda: ops.norm_bwd(xb, x, w, st, torch.float32, dw=dw, dres=dres)), 7 * u)
    report("LN bwd fp32 dy, bf16 x -> bf16", timeit(lambda: ops.norm_bwd(x, xb, w, st, torch.bfloat16, dw=dw)), 4 * u)
    report("cast fp32->bf16", timeit(lambda: ops.cast_bf16(x)), 3 * u)
    ids_e = torch.randint(0, 1024, (256, 257), device=dev)
    ids_e[torch.rand(256, 257, device=dev) < 0.5] = 2024
    dword, dpos = torch.zeros(2025, H, device=dev), torch.zeros(257, H, device=dev)
    report("embed bwd (half of the tokens = mask id)", timeit(lambda: ops.embed_bwd(ids_e, x, dword, dpos)), 2 * u)
if want("attn"):
    qkv = bf(T, 3 * H)
    scale = 0.125
    o, lse = ops.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, nh, S, S, scale)
    do = bf(T, H)
    dqkv = torch.empty_like(qkv)
    fl = 4 * S * S * 64 * B * nh / 1e9
    report("attn fwd", timeit(lambda: ops.attn_fwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], B, nh, S, S, scale)), gflop=fl)
    report("attn bwd", timeit(lambda: ops.attn_bwd(qkv[:, :H], qkv[:, H:2 * H], qkv[:, 2 * H:], o, do, lse, dqkv[:, :H],
                                                   dqkv[:, H:2 * H], dqkv[:, 2 * H:], B, nh, S, S, scale)), gflop=2.5 * fl)
if want("gemm"):
    x = bf(T, H)
    for name, N, K in [("qkv", 3 * H, H), ("out", H, H), ("wi", 2 * I, H), ("wo", H, I), ("logits", 2048, H)]:
        w = bf(N, K)
        xin = bf(T, K)
        fl = 2 * T * N * K / 1e9
        report(f"gemm fwd {name} [T,{K}]x[{N},{K}]", timeit(lambda: ops.linear_fwd(xin, w)), gflop=fl)
        dyy = bf(T, N)
        report(f"gemm dgrad {name}", timeit(lambda: ops.linear_dgrad(dyy, w)), gflop=fl)
        dwt = torch.zeros(N, K, device=dev)
        report(f"gemm wgrad {name}", timeit(lambda: ops.linear_wgrad(dyy, xin, dwt)), gflop=fl)
    res = torch.randn(T, H, device=dev)
    w = bf(H, I)
    xin = bf(T, I)
    report("gemm wo + residual (fp32 out)", timeit(lambda: ops.linear_fwd(xin, w, res=res)), gflop=2 * T * H * I / 1e9)
if want("resgemm"):  # the two residual-epilogue GEMMs of a layer (attention out-projection, FFN wo)
    res = torch.randn(T, H, device=dev)
    for name, K in [("attn out", H), ("wo", I)]:
        w = bf(H, K)
        xin = bf(T, K)
        report(f"gemm {name} + residual (fp32 out)", timeit(lambda: ops.linear_fwd(xin, w, res=res), n=20), gbytes=(T * K * 2 + 2 * T * H * 4) / 1e6,
               gflop=2 * T * H * K / 1e9)
