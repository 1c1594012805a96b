#!/bin/bash
# ncu --set full with source correlation: one launch of each attention backward kernel and of the two GLU+LN kernels
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ncu_bwd.log) 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd" -s 4 -c 2 -f -o gpurun_out/r2_attn_bwd python tools/bench_kernels.py attn > /dev/null 2>&1
ls -la gpurun_out/r2_attn_bwd.ncu-rep
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"glu_norm_bwd" -s 16 -c 1 -f -o gpurun_out/r2_glu_bwd python tools/bench_kernels.py glu > /dev/null 2>&1
ls -la gpurun_out/r2_glu_bwd.ncu-rep
echo "=== DONE"
