#!/usr/bin/env python
"""Launches every named kernel of the hot path ONCE at its BASELINE-config shape between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/r2_kernels python tools/profile_kernels.py
The same kernel list bench.py times with CUDA events (bench.kernel_rooflines -> "roofline_kernels" in the bench line)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)


def once(fn, n=1, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return 1.0


bench._timeit = once
peak_tf, peak_hbm, _ = bench.measured_peaks()
rows = bench.kernel_rooflines(dev, peak_tf, peak_hbm)
print(f"profiled {len(rows)} kernel groups")
