#!/bin/bash
# round-2 quick loop: build check, GPU tests (all failures shown), bench line
mkdir -p gpurun_out
LOG=gpurun_out/${LOG_NAME:-r2_quick}.log
exec > >(tee $LOG) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1800 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/${LOG_NAME:-r2_quick}_pytest_full.log 2>&1
grep -v "^$" gpurun_out/${LOG_NAME:-r2_quick}_pytest_full.log | grep -E "^(FAILED|ERROR|E  |tests/.*(Error|assert)|[0-9]+ (passed|failed))|rel-L2|agree|reference train script|single-pass" | cut -c1-400 | tail -${TAIL:-80}
if [ -z "$NO_BENCH" ]; then echo "=== BENCH"; timeout 900 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} 2>&1 | tail -3; fi
echo "=== DONE"
