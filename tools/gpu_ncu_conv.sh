#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/ncu_conv.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
python tools/bench_kernels.py glu 2>/dev/null | tail -6
VQ_BATCH=16 timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"conv_tc_kernel|gn_apply_silu_split8" -s 1 -c 5 -f -o gpurun_out/prof_conv_tc python tools/profile_vqgan.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
echo "=== DONE"
