"""Aggregate an ncu source-page CSV (--print-source sass,cuda) by CUDA source line.
usage: ncu -i rep --page source --print-source sass,cuda --csv -k regex:X ... | python scripts/ncu_lines.py [top]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 25
rows = list(csv.reader(sys.stdin))
cur_file = ""
agg = {}
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"):
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iT = hdr.index("Thread Instructions Executed")
        continue
    if hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:90])
            agg[key] = (int(r[iI]), int(r[iS]), int(r[iT]))
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot_i:,}  samples {tot_s:,}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    eff = v[2] / v[0] if v[0] else 0
    print(f"{100*v[0]/tot_i:5.1f}% inst {100*v[1]/max(1,tot_s):5.1f}% smp  thr/inst {eff:4.1f}  {k[0]}:{k[1]}  {k[2]}")
