#!/bin/bash
# kernel tests + A/B-able bench (pass env through)
mkdir -p gpurun_out
exec > >(tee gpurun_out/ab.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | grep -v "^$" | cut -c1-300 | tail -8
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-full-step 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"; done
python tools/bench_kernels.py norm 2>/dev/null | tail -8
echo "=== DONE"
