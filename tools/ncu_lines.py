#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep captured with
--import-source on (run here, no GPU needed):  python tools/ncu_lines.py <rep> [min_pct]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) == len(hdr):
        try:
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        samples = r[hdr.index("# Samples")]
        thr = r[hdr.index("Avg. Threads Executed")]
        lines.append((fname, int(r[0]), r[1].strip(), inst, int(samples) if samples.isdigit() else 0, thr))
tot = sum(l[3] for l in lines) or 1
tots = sum(l[4] for l in lines) or 1
print(f"total warp instructions {tot}, samples {tots}")
for f, n, src, inst, smp, thr in lines:
    if 100.0 * inst / tot >= min_pct or 100.0 * smp / tots >= min_pct:
        print(f"{f}:{n:5d} {100.0 * inst / tot:6.2f}% inst {100.0 * smp / tots:6.2f}% smp thr {thr:>5s} | {src[:110]}")
