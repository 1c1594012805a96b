#!/bin/bash
# last validation of the round: full GPU suite and the default bench line on the final tree
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final3.log) 2>&1
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_final3_pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^$" gpurun_out/r2_final3_pytest.log | grep -E "^(FAILED|ERROR|E  |[0-9]+ (passed|failed))" | cut -c1-300 | tail -20
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_final3_bench.json") if l.startswith("{")][-1])
print("step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s | e2e", round(d["e2e"]["value"]), "| gemm frac", round(d["roofline"]["frac"], 3), "| clocks", d.get("clocks"))
for k in ("decode_steps_per_s", "cpu_baseline"):
    print(" ", k, json.dumps(d[k])[:300])
for r in d.get("roofline_kernels", []):
    print("   %-88s %8.1f us  frac %.3f" % (r["kernel"][:88], r["us"], r["frac"]))
PY
echo "=== DONE"
