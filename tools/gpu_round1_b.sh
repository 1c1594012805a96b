#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_b.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
echo "=== ALL GPU TESTS"; timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | tail -60
echo "=== SMOKE"; timeout 300 python __graft_entry__.py --smoke
echo "=== BENCH mma"; MUSE_B200_GEMM=mma timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline
echo "=== BENCH tcgen05"; timeout 900 python bench.py --steps 10 --warmup 3
echo "=== DONE"
