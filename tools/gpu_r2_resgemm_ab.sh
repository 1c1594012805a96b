#!/bin/bash
# A/B of the next-tile residual L2 prefetch in the GEMM's residual epilogue (same box, same call)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_resgemm_ab.log) 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "gemm" 2>&1 | tail -2
for v in 1 0 1 0; do echo "=== MUSE_B200_RES_PREFETCH=$v"; MUSE_B200_RES_PREFETCH=$v timeout 200 python tools/bench_kernels.py resgemm 2>&1 | grep -E "residual"; done
echo "=== DONE"
