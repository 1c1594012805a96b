#!/bin/bash
# focused dev cycle: build, a pytest subset (PYTEST_ARGS), the bench line with per-kernel rooflines (compact print)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_cycle.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 1500 python -m pytest ${PYTEST_ARGS:-tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_uvit_v2_gpu.py} -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_cycle_pytest.log 2>&1
grep -v "^$" gpurun_out/r2_cycle_pytest.log | grep -E "^(FAILED|ERROR|E  |tests/.*(Error|assert)|[0-9]+ (passed|failed))|rel-L2|stages:" | cut -c1-300 | tail -${TAIL:-40}
if [ -z "$NO_BENCH" ]; then
  timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-full-step ${BENCH_ARGS} 2>/dev/null | grep -E '^\{' > gpurun_out/r2_cycle_bench.json
  python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_cycle_bench.json").readline())
print("step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s; e2e", round(d["e2e"]["value"]), "; gemm frac", round(d["roofline"]["frac"], 3))
for k in ("decode_steps_per_s", "t2i_pipeline_latency"):
    if k in d:
        print(k, json.dumps(d[k])[:900])
for r in d.get("roofline_kernels", []):
    print("  %-88s %8.1f us  frac %.3f" % (r["kernel"][:88], r["us"], r["frac"]))
PY
fi
echo "=== DONE"
