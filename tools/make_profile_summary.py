#!/usr/bin/env python
"""Turns an `ncu --set full` report (exported with --page raw --csv) into the short per-kernel table kept in profiles/."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_%active"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# source: {rep} (ncu --set full --clock-control none); one line per captured launch\n")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("muse::<unnamed>::", "")
            parts = [name[:60]]
            for k, label in KEYS:
                if k in idx:
                    parts.append(f"{label}={r[idx[k]]}{units[idx[k]].replace('byte', 'B') if units[idx[k]] not in ('', '%') else units[idx[k]]}")
            f.write("  ".join(parts) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
