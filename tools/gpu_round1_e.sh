#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_e.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-300 | tail -25
echo "=== LAUNCH LIST (ncu, one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py
echo "=== BENCH"; timeout 900 python bench.py --steps 10 --warmup 3
echo "=== NCU FULL: gemm + attention bwd"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tcgen05_kernel|attn_bwd_dq_kernel|attn_fwd_kernel" -s 10 -c 12 -o gpurun_out/prof_r1 python tools/profile_step.py
echo "=== DONE"
