#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [launch_skip] [topN] [smp|inst]   (sort key, default samples)"""
import csv, subprocess, sys, collections, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"; top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
key = 1 if (len(sys.argv) > 5 and sys.argv[5] == "inst") else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = ""; hdr = None; agg = collections.defaultdict(lambda: [0, 0, 0, 0, ""])
cur = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed"); iL = hdr.index("stall_long_sb"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0]: cur = (fname, int(r[0])); agg[cur][4] = r[1].strip()[:110]
    if cur is None or not r[2]: continue
    def num(x):
        try: return int(float(x))
        except ValueError: return 0
    a = agg[cur]; a[0] += num(r[iS]); a[1] += num(r[iI]); a[2] += num(r[iT]); a[3] += num(r[iL])
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][key])[:top]:
    eff = a[2] / a[1] if a[1] else 0
    print(f"{a[0]/tot*100:5.1f}% smp {a[1]/toti*100:5.1f}% inst  thr/inst {eff:4.1f}  longsb {a[3]/tot*100:4.1f}%  {k[0]}:{k[1]}  {a[4]}")
