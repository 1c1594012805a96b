#!/bin/bash
# compute-sanitizer pass over the kernel-level GPU tests (the reference has no race / memory checking; the
# CUDA path should).  Never run in round 2 (the GPU budget went to parity, profiles and benches) -- second call to make when a
# B200 is available again, after tools/gpu_pending.sh.  memcheck over every kernel test, racecheck + synccheck over the
# shared-memory-heavy ones (norms, GLU, sampling, VQ, attention, GEMM at their smallest shapes).  Sanitized runs are 10-100x
# slower: shapes are limited with -k, and each tool gets its own timeout.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_sanitizer.sh'
mkdir -p gpurun_out
exec > >(tee gpurun_out/sanitizer.log) 2>&1
PY="python -m pytest -m gpu -x -q --tb=short -p no:cacheprovider"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 $PY tests/test_kernels_gpu.py \
    -k "embed or norm or glu or cross_entropy or attention" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 $PY tests/test_kernels_gpu.py \
    -k "norm or glu or cross_entropy" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 20 $PY tests/test_kernels_gpu.py \
    -k "attention or gemm" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -5 gpurun_out/sanitizer_synccheck.log
echo "=== DONE"
