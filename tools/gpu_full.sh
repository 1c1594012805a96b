#!/bin/bash
# full loop: all GPU tests, launch list, bench (incl. full step), aux benches, torch-eager-on-GPU reference
bash tools/gpu_quick.sh
exec > >(tee gpurun_out/full_extra.log) 2>&1
echo "=== AUX"; timeout 900 python tools/bench_aux.py generate2 vqgan taming uvit
echo "=== EAGER ORACLE ON GPU"; timeout 600 python bench.py --impl reference --ref-device cuda --steps 5 --warmup 2
echo "=== C4 / UVIT train"; timeout 300 python tools/bench_c4.py --steps 4 2>&1 | grep "^{"; timeout 300 python tools/bench_c4.py --model uvit --steps 4 2>&1 | grep "^{"
echo "=== SMOKE"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "=== DONE2"
