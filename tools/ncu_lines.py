#!/usr/bin/env python
"""Aggregate an ncu report's per-line metrics: ncu_lines.py REPORT KERNEL_REGEX [top]
(uses `ncu --page source --print-source cuda,sass --csv`)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
cur, hdr, rows = None, None, []
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0] != "":
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        rows.append((cur, int(r[0]), r[1], d))
def f(d, k):
    try: return float(d[k])
    except Exception: return 0.0
K, S = "Instructions Executed", "# Samples"
byfile = {}
for fn, ln, src, d in rows:
    a = byfile.setdefault(fn, [0, 0]); a[0] += f(d, K); a[1] += f(d, S)
tot = sum(a[0] for a in byfile.values()); tots = sum(a[1] for a in byfile.values())
print("total warp-instr %.0f samples %.0f" % (tot, tots))
for fn, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print("  %-22s inst %5.1f%%  samples %5.1f%%" % (fn, 100 * a[0] / max(tot, 1), 100 * a[1] / max(tots, 1)))
for fn, ln, src, d in sorted(rows, key=lambda x: -f(x[3], K if len(sys.argv) > 4 else S))[:top]:
    print("%-18s %5d inst %5.2f%% smp %5.2f%% thr/inst %4.1f | %s" % (fn, ln, 100 * f(d, K) / max(tot, 1), 100 * f(d, S) / max(tots, 1), f(d, "Avg. Threads Executed"), src.strip()[:80]))
