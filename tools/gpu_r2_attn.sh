#!/bin/bash
# attention bring-up: kernel tests first (short timeout), kernel timings, then the model tests and a bench line
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_attn.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | grep -v "^$" | cut -c1-300 | tail -40
echo "=== kernel timings"
timeout 240 python tools/bench_kernels.py attn 2>&1 | tail -5
echo "=== model tests"
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_uvit_v2_gpu.py tests/test_sampling_vqgan_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|^E  |rel-L2" | cut -c1-300 | tail -30
echo "=== BENCH"; timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-full-step 2>&1 | tail -1 | cut -c1-600
echo "=== DONE"
