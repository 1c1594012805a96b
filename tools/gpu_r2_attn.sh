#!/bin/bash
# attention bring-up: kernel tests first (short timeout), kernel timings, then the model tests and a bench line
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_attn.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider -k "attention" 2>&1 | grep -v "^$" | cut -c1-300 | tail -40
echo "=== kernel timings"
timeout 240 python tools/bench_kernels.py attn 2>&1 | tail -5
echo "=== ncu stall summary (bwd kernels)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd" -s 4 -c 2 -f -o gpurun_out/r2_attn_bwd python tools/bench_kernels.py attn > /dev/null 2>&1
ncu -i gpurun_out/r2_attn_bwd.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active']
idx = [hdr.index(w) for w in want if w in hdr]
stall = [i for i, h in enumerate(hdr) if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    print(' | '.join(r[i][:50] for i in idx))
    for i in sorted(stall, key=lambda i: -float(r[i] or 0))[:6]:
        print('    ', hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i])
"
echo "=== model tests"
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | grep -E "passed|failed|FAILED|^E  |rel-L2" | cut -c1-300 | tail -30
echo "=== BENCH"; timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --no-full-step 2>&1 | tail -1 | cut -c1-600
echo "=== DONE"
