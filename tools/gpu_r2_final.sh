#!/bin/bash
# Final round-2 validation in ONE gpurun call (the prebuilt .so files travel with the snapshot; nothing is compiled here):
#   1. full `pytest -m gpu` with the library defaults
#   2. the same suite with programmatic dependent launch forced on (MUSE_B200_PDL=1)
#   3. bench line with --pdl 1 (secondary metrics included, CPU baseline / from-pixels step skipped)
#   4. the full default bench line (the evidence copied to profiles/)
#   5. ncu --set full of the three attention kernels (stall reasons), if time is left
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_final.log) 2>&1
T0=$(date +%s)
el() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.sm,power.limit --format=csv,noheader
ls -la open_muse_b200/libmuse_b200.so tests/xcheck/libmuse_b200_xcheck.so oracle/_build/libvq_oracle.so

summ() {  # failures + the summary line of a pytest log
  grep -v "^$" "$1" | grep -E "^(FAILED|ERROR|E  |tests/.*(Error|assert)|[0-9]+ (passed|failed))" | cut -c1-300 | tail -${2:-25}
}

PDLTEST=tests/test_model_gpu.py::test_programmatic_dependent_launch_is_bit_identical
el "pytest (library defaults; the PDL on/off test runs in its own process below)"
timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --deselect $PDLTEST > gpurun_out/r2_final_pytest_default.log 2>&1
echo "rc=$?"; summ gpurun_out/r2_final_pytest_default.log

el "pytest PDL on/off bit-identity test"
timeout 300 python -m pytest $PDLTEST -q --tb=long -p no:cacheprovider > gpurun_out/r2_final_pytest_pdltest.log 2>&1
echo "rc=$?"; summ gpurun_out/r2_final_pytest_pdltest.log 60

el "pytest (MUSE_B200_PDL=1)"
MUSE_B200_PDL=1 timeout 700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/r2_final_pytest_pdl1.log 2>&1
echo "rc=$?"; summ gpurun_out/r2_final_pytest_pdl1.log

show() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
except Exception as e:
    print("no bench line in", sys.argv[1], e); sys.exit(0)
print("pdl", d["config"].get("pdl"), "| step", round(d["ms_per_step"], 3), "ms", round(d["value"]), "img/s | e2e", round(d["e2e"]["value"]),
      "| gemm frac", round(d["roofline"]["frac"], 3), "| clocks", d.get("clocks"))
for k in ("decode_steps_per_s", "t2i_pipeline_latency", "vqgan_roundtrip", "full_step_incl_vq_encode", "torch_eager_same_gpu", "cpu_baseline"):
    if d.get(k):
        print(" ", k, json.dumps(d[k])[:700])
for r in d.get("roofline_kernels", []):
    print("   %-88s %8.1f us  frac %.3f" % (r["kernel"][:88], r["us"], r["frac"]))
PY
}

el "bench --pdl 1"
timeout 600 python bench.py --steps 10 --warmup 3 --pdl 1 --no-cpu-baseline --no-full-step > gpurun_out/r2_final_bench_pdl1.json 2> gpurun_out/r2_final_bench_pdl1.err
echo "rc=$?"; tail -3 gpurun_out/r2_final_bench_pdl1.err | cut -c1-300; show gpurun_out/r2_final_bench_pdl1.json

el "bench (defaults, full line)"
timeout 700 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_final_bench_default.json 2> gpurun_out/r2_final_bench_default.err
echo "rc=$?"; tail -3 gpurun_out/r2_final_bench_default.err | cut -c1-300; show gpurun_out/r2_final_bench_default.json

el "ncu attention kernels"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:"attn_" -s 6 -c 3 -f -o gpurun_out/r2_final_attn python tools/bench_kernels.py attn > /dev/null 2>&1
ls -la gpurun_out/r2_final_attn.ncu-rep
ncu -i gpurun_out/r2_final_attn.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
if len(rows) < 3: sys.exit(0)
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active']
idx = [hdr.index(w) for w in want if w in hdr]
stall = [i for i, h in enumerate(hdr) if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    print(' | '.join(r[i][:60] for i in idx))
    for i in sorted(stall, key=lambda i: -float(r[i] or 0))[:8]:
        print('    ', hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i])
"
el "DONE"
