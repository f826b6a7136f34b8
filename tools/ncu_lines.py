"""Development aid: per-source-line warp-stall samples of the kernels in an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [min_percent]      (one section per source file / inlined function)"""
import csv, subprocess, sys
rep = sys.argv[1]; minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
sections = []          # (title, [(line, samples, instructions, source)])
title = "?"; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "Function Name", "Kernel Name") and len(r) > 1: title = r[1]; continue
    if r[0] == "Line No":
        hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); cur = (title, {}); sections.append(cur); continue
    if cur is None or len(r) <= iS or r[0] == "": continue
    try: ln, s, n = int(r[0]), int(r[iS]), int(r[iI])
    except ValueError: continue
    if ln not in cur[1]: cur[1][ln] = (s, n, r[1][:120])
tot = sum(v[0] for _, d in sections for v in d.values())
print("total samples", tot)
for name, d in sections:
    fs = sum(v[0] for v in d.values())
    if fs == 0 or 100.0 * fs / tot < minp: continue
    print(f"== {name}: {100.0*fs/tot:.1f}% of the samples")
    for ln in sorted(d):
        s, n, src = d[ln]
        if 100.0 * s / tot >= minp: print(f"{ln:4d} {100.0*s/tot:5.1f}% {n:10d}  {src}")
