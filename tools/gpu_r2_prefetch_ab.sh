#!/bin/bash
# A/B of the L2 prefetch hints in the persistent norm / GLU backward kernels (same box, same call)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_prefetch_ab2.log) 2>&1
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "norm or glu" 2>&1 | tail -2
for v in 1 0 1 0; do echo "=== MUSE_B200_ROW_PREFETCH=$v"; MUSE_B200_ROW_PREFETCH=$v timeout 200 python tools/bench_kernels.py glu 2>&1 | grep -E "saved y"; done
echo "=== bench step (prefetch on / off / on / off)"
for v in 1 0 1 0; do MUSE_B200_ROW_PREFETCH=$v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-full-step --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('prefetch $v: step', round(d['ms_per_step'],3), 'ms', round(d['value']), 'img/s clocks', d['clocks']['sm_mhz'])"; done
echo "=== DONE"
