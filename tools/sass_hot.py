#!/usr/bin/env python
"""Segment the SASS page of an ncu report by executed-count to find the hot regions.
usage: ncu -i X.ncu-rep --page source --csv > /tmp/s.csv; python tools/sass_hot.py /tmp/s.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    body.append(r)
ia, it, isrc, isamp = (hdr.index(x) for x in ("Instructions Executed", "Avg. Threads Executed", "Source", "# Samples"))
tot = sum(int(r[ia]) for r in body)
tsamp = sum(int(r[isamp]) for r in body)
print("total warp inst", tot, "n sass", len(body), "samples", tsamp)
prev, start, seg = None, 0, []
for i, r in enumerate(body):
    key = (int(r[ia]), r[it])
    if key != prev:
        if prev is not None:
            seg.append((start, i - 1, prev))
        start, prev = i, key
seg.append((start, len(body) - 1, prev))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
for s, e, (cnt, thr) in seg:
    n = e - s + 1
    samples = sum(int(body[k][isamp]) for k in range(s, e + 1))
    if 100.0 * cnt * n / tot >= thresh or 100.0 * samples / max(tsamp, 1) >= thresh:
        print(f"sass[{s:4d}-{e:4d}] n={n:4d} exec={cnt:8d} thr={thr:>5s} inst={100*cnt*n/tot:5.1f}% "
              f"samples={100*samples/max(tsamp,1):5.1f}%  {body[s][isrc].strip()[:60]}")
