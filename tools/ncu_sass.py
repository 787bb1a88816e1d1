#!/usr/bin/env python
"""Top SASS instructions by PC samples from an .ncu-rep (source page). usage: ncu_sass.py rep kernel_regex [launch_skip] [topN]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"; top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; items = []
for r in rows:
    if len(r) > 6 and "Source" in r and "# Samples" in r: hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: s = int(float(d["# Samples"]))
    except Exception: continue
    items.append((s, d))
tot = sum(s for s, _ in items) or 1
print("total samples", tot, "instructions", len(items))
stallcols = [c for c in hdr if c.startswith("stall_")]
for i, (s, d) in enumerate(items):
    d["_idx"] = i
for s, d in sorted(items, key=lambda x: -x[0])[:top]:
    st = sorted(((c[6:], int(float(d[c] or 0))) for c in stallcols), key=lambda x: -x[1])[:3]
    ie = float(d.get("Instructions Executed") or 0); te = float(d.get("Thread Instructions Executed") or 0)
    print(f"{s/tot*100:5.2f}%  #{d['_idx']:4d} thr/inst {te/ie if ie else 0:4.1f}  {d['Source'][:70]:70s} " + " ".join(f"{k}={v}" for k, v in st if v))
