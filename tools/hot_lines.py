#!/usr/bin/env python
"""Hot source lines of one kernel launch in an ncu report captured with --import-source on.
usage: tools/hot_lines.py report.ncu-rep <launch-skip> [top]"""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, agg, fn = None, [], None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == 'Function Name': fn = r[1]; continue
    if len(r) > 7 and r[0].isdigit():
        try: agg.append((int(r[6] or 0), int(r[7] or 0), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError: pass
tot = sum(a[0] for a in agg) or 1; toti = sum(a[1] for a in agg) or 1
print(fn[:110]); print("samples", tot, "warp-instructions", toti)
for s, i, f, l, src in sorted(agg, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins  {f}:{l}  {src}")
