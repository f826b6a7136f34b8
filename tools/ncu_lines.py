"""Development aid: per-source-line warp-stall samples of the first kernel in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [min_percent]"""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[h]; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed")
seen = {}; 
for r in rows[h + 1:]:
    if len(r) <= iS or r[0] == "": continue
    try: ln = int(r[0]); s = int(r[iS]); n = int(r[iI])
    except ValueError: continue
    if ln in seen: break          # second launch starts
    seen[ln] = (s, n, r[1][:120])
tot = sum(v[0] for v in seen.values())
print("total samples", tot)
for ln in sorted(seen):
    s, n, src = seen[ln]
    if 100.0 * s / tot >= minp: print(f"{ln:4d} {100.0*s/tot:5.1f}% {n:9d}  {src}")
