#!/usr/bin/env python
"""Top stall locations (SASS) of one launch in an .ncu-rep: ncu_hot.py rep launch_index [min_pct]"""
import csv, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    try:
        data.append((int(r[i_s]), int(r[i_ex]), r[i_src]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
acc = 0
for n, (s, ex, src) in enumerate(data):
    acc += s
    if s > tot * minp / 100:
        print("%5d %5.1f%% cum %5.1f%% ex=%9d  %s" % (n, 100 * s / tot, 100 * acc / tot, ex, src[:100]))
