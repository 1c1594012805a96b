#!/bin/bash
mkdir -p gpurun_out
exec > >(tee gpurun_out/round1_d.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 python -m pytest tests -m gpu -q -s --tb=short -p no:cacheprovider 2>&1 | grep -v "^$" | cut -c1-300 | tail -60
echo "=== LAUNCH LIST (ncu, one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py
echo "=== BENCH"; timeout 900 python bench.py --steps 10 --warmup 3
echo "=== DONE"
