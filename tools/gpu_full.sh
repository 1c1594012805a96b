#!/bin/bash
# full loop: all GPU tests, launch list, bench (incl. full step), aux benches, torch-eager-on-GPU reference
bash tools/gpu_quick.sh
exec > >(tee gpurun_out/full_extra.log) 2>&1
echo "=== AUX"; timeout 900 python tools/bench_aux.py generate2 vqgan taming uvit
echo "=== EAGER ORACLE ON GPU"; timeout 600 python bench.py --impl reference --ref-device cuda --steps 5 --warmup 2
echo "=== DONE2"
