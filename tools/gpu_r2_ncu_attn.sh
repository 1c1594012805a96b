#!/bin/bash
# ncu --set full with source correlation for the attention kernels (one launch each)
mkdir -p gpurun_out
exec > >(tee gpurun_out/r2_ncu_attn.log) 2>&1
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_" -s 6 -c 3 -f -o gpurun_out/r2_attn python tools/bench_kernels.py attn > /dev/null 2>&1
ls -la gpurun_out/r2_attn.ncu-rep
ncu -i gpurun_out/r2_attn.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed_pipe_xu.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print(' | '.join(r[i][:60] for i in idx))
stall = [i for i, h in enumerate(hdr) if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    print(r[hdr.index('Kernel Name')][:50])
    for i in sorted(stall, key=lambda i: -float(r[i] or 0))[:8]:
        print('    ', hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i])
"
echo "=== DONE"
