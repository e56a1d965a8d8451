"""Turn the raw ncu outputs in gpurun_out/ into the small, committed summaries under profiles/.
    python scripts/summarise_profiles.py <launches.csv> <full.ncu-rep> <round tag>"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
out_dir = os.path.join(ROOT, "profiles")

# ---- launch list: share of the step per kernel ----
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("gicp::", "")
    name = name.split("<")[0] if not name.startswith("cub") and not name.startswith("at::") else name.split("<")[0]
    agg[name][0] += 1
    agg[name][1] += float(r[iv].replace(",", "")) / 1e6   # ns -> ms
total = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"launches_{tag}.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}): `bench.py --pairs 512 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e`\n\n"
            "`ncu --metrics gpu__time_duration.sum --clock-control none` - per-launch times are cold-cache and\n"
            "serialised: read the SHARES, not the absolutes.  All launches of the library's kernels and of CUB (4 warm-up\n"
            "+ 1 timed + 1 profiled step; `--kernel-name regex:` keeps the torch kernels that generate the synthetic\n"
            "clouds out of the list).\n\n"
            "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {100 * v[1] / total:.1f}% |\n")
print("launch list:", len(agg), "kernels, total", round(total, 1), "ms")

# ---- full capture: one row per profiled launch ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
H = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
idx = [H.index(w) for w in want]
traffic = {}
with open(os.path.join(out_dir, f"ncu_full_{tag}.md"), "w") as f:
    f.write(f"# ncu --set full ({tag}): `scripts/profile_step.py 64` (64 pairs x 32768 points, first launches)\n\n"
            "Units as printed by ncu (time ms, bytes MB).  Launch order: k-NN(target), k-NN(source), then\n"
            "correspond/accumulate of outer iterations 0, 1, 2.\n\n| " + " | ".join(w.split(".")[0] for w in want) + " |\n|" + "---|" * len(want) + "\n")
    for r in rr[2:]:
        vals = [r[i] for i in idx]
        vals[0] = vals[0].split("(")[0].replace("void ", "")
        f.write("| " + " | ".join(v[:60] for v in vals) + " |\n")
        name = vals[0].split("<")[0]
        byt = (float(r[H.index("dram__bytes_read.sum")]) + float(r[H.index("dram__bytes_write.sum")])) * 1e6
        traffic.setdefault(name, []).append(byt)
json.dump({"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of "
                   "scripts/profile_step.py 64 (2,097,152 points per cloud side); bytes", "launches": traffic,
           "points_per_launch": 64 * 32768}, open(os.path.join(out_dir, f"traffic_{tag}.json"), "w"), indent=1)
print("full capture:", len(rr) - 2, "launches")
