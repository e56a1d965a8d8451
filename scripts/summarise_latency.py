"""profiles/ncu_latency_<tag>.md from the two `ncu --set full` captures of the latency path
(scripts/latency_breakdown.py 360 / 90):  python scripts/summarise_latency.py <rep360> <rep90> <tag>"""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep360, rep90, tag = sys.argv[1], sys.argv[2], sys.argv[3]
COLS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
lines = []
for rep, rays in ((rep360, 360), (rep90, 90)):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].replace("void ", "").split("(")[0]
        vals = []
        for c in COLS:
            v = r[h.index(c)]
            try:
                v = f"{float(v.replace(',', '')):.2f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            vals.append(f"{v} {units[h.index(c)]}".strip())
        lines.append(f"| {name} | " + " | ".join(vals) + f" | {rays} |")
with open(os.path.join(ROOT, "profiles", f"ncu_latency_{tag}.md"), "w") as f:
    f.write(f"# ncu --set full ({tag}) of the latency path: `scripts/latency_breakdown.py 360` and `... 90` (one pair of scans)\n\n"
            "Launches 60+ of the process (warm).  `register_loop_kernel` is ONE block (one pair) running every outer\n"
            "iteration, including the initialisation of the pair; `small_grid_kernel` / `knn_*` are one block (or 3-12\n"
            "warps) per cloud, the two clouds of a pair side by side on two streams (`gicpSetPair`).\n\n")
    f.write("| kernel | " + " | ".join(c.split(".")[0] for c in COLS) + " | rays |\n")
    f.write("|" + "---|" * (len(COLS) + 2) + "\n")
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
