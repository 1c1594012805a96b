#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/ncu_glu.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"glu_norm_fwd_kernel|glu_norm_bwd_kernel" -s 4 -c 2 -f -o gpurun_out/prof_glu python tools/bench_kernels.py glu > /dev/null 2>&1
ls -la gpurun_out/prof_glu.ncu-rep
echo "=== DONE"
