#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an `ncu --page source --csv` export (optionally gzipped)."""
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    h = rows[1]
    si, src, ex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    data = []
    for k, r in enumerate(rows[2:]):
        try:
            data.append((float(r[si]), float(r[ex]), k, r[src].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1
    totex = sum(d[1] for d in data)
    print("# %s: %d SASS lines, %.0f samples, %.3g warp instructions" % (path, len(data), tot, totex))
    for s, e, k, text in sorted(data, key=lambda d: -d[0])[:top]:
        print("%6.2f%%  line %5d  exec %12.0f  %s" % (100 * s / tot, k, e, text[:100]))


if __name__ == "__main__":
    main()
