#!/bin/bash
# A/B of scheduling options on one box: wgrad side stream on/off, fused optimizer
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ab.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
for args in "--no-wgrad-overlap" "" "--optimizer fused" "--no-cuda-graph" "--no-cuda-graph --no-wgrad-overlap"; do
  echo "=== bench $args"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu-baseline --no-full-step $args 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('   ', round(d['value']), 'img/s', round(d['ms_per_step'], 3), 'ms/step; e2e', round(d['e2e']['value']), '; gemm frac', round(d['roofline']['frac'], 3), 'share', round(d['roofline']['gemm_share_of_step'], 3), d['config'].get('cuda_graph'), d['config'].get('optimizer'), d['config'].get('wgrad_side_stream'))
"
done
echo "=== DONE"
