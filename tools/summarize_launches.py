#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a
markdown table (per kernel: launches, total, average, share of the GPU time).

    python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_....md
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r[4]).replace("void ", "")[:80]
        a = agg.setdefault((name, r[7], r[8]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[14])
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k[0].startswith("mgb::"))
    print("# %s\n" % title)
    print("Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes.\n")
    print("%d launches, %.3f ms of GPU time, %.1f%% of it in this library's kernels (`mgb::`).\n"
          % (len(rows), tot / 1e6, 100 * ours / tot))
    print("| kernel | block | grid | launches | total ms | avg ms | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        if a[1] / tot < 0.0005:
            continue
        print("| `%s` | %s | %s | %d | %.3f | %.3f | %.1f%% |"
              % (k[0], k[1], k[2], a[0], a[1] / 1e6, a[1] / a[0] / 1e6, 100 * a[1] / tot))


if __name__ == "__main__":
    main()
