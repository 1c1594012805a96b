#!/bin/bash
# first GPU bring-up call: diagnostics -> kernel tests -> model tests -> short bench (both GEMM backends)
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_a.log) 2>&1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
echo "=== DIAG"; timeout 1500 python tools/gpu_diag.py
echo "=== KERNEL TESTS (non-tcgen05)"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "not tcgen05" -p no:cacheprovider 2>&1 | tail -40
echo "=== KERNEL TESTS (tcgen05)"; timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "tcgen05" -p no:cacheprovider 2>&1 | tail -40
echo "=== MODEL TESTS (mma backend)"; MUSE_B200_GEMM=mma timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "not tcgen05" -p no:cacheprovider 2>&1 | tail -40
echo "=== MODEL TESTS (tcgen05 backend)"; timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -s -k "not mma" -p no:cacheprovider 2>&1 | tail -40
echo "=== SMOKE"; timeout 300 python __graft_entry__.py --smoke
echo "=== BENCH mma"; MUSE_B200_GEMM=mma timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
echo "=== BENCH tcgen05"; timeout 600 python bench.py --steps 5 --warmup 3
echo "=== DONE"
