#!/usr/bin/env python
"""Compact per-source-line instruction counts from an ncu report:
   python tools/ncu_lines.py report.ncu-rep [file-substring] [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else "render.cu"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fn = None; cur = None; H = None; acc = {}
for r in csv.reader(io.StringIO(txt)):
    if not r: continue
    if r[0] == "Function Name": fn = r[1].split("(")[0].split("::")[-1]; continue
    if r[0] == "File Path" or r[0] == "File Name": cur = r[1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or cur is None or sub not in cur: continue
    try:
        ln = int(r[0]); ie = int(r[H.index("Instructions Executed")] or 0); smp = int(r[H.index("# Samples")] or 0)
    except Exception: continue
    acc.setdefault(fn, []).append((ie, smp, ln, r[1].strip()[:90]))
for fn, rows in acc.items():
    tot = sum(x[0] for x in rows) or 1; ts = sum(x[1] for x in rows) or 1
    print("== %s  (instr in %s: %.1fM)" % (fn, sub, tot / 1e6))
    for ie, smp, ln, src in sorted(rows, reverse=True)[:top]:
        print("%5.1f%% inst %5.1f%% smp  L%-4d %s" % (100 * ie / tot, 100 * smp / ts, ln, src))
if len(sys.argv) > 4:
    # python tools/ncu_lines.py rep file top "fn:lo-hi,lo-hi,..."
    fn, spec = sys.argv[4].split(":")
    rows = acc[fn]; tot = sum(x[0] for x in rows)
    for rg in spec.split(","):
        lo, hi = map(int, rg.split("-"))
        s = sum(x[0] for x in rows if lo <= x[2] <= hi)
        print("%s L%d-%d: %.1fM instr (%.1f%%)" % (fn, lo, hi, s / 1e6, 100 * s / tot))
