#!/bin/bash
# quick loop: kernel+model tests, one-step launch list, bench
mkdir -p gpurun_out
exec > >(tee gpurun_out/quick.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS} 2>&1 | grep -v "^$" | cut -c1-300 | tail -25
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv python tools/profile_step.py > /dev/null
echo "=== BENCH"; timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS}
echo "=== DONE"
