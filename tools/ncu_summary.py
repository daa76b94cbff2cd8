#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean us, share.
    python tools/ncu_summary.py gpurun_out/launches.csv [--skip-peaks] > profiles/rNN_launches_summary.txt"""
import collections
import csv
import io
import sys


def main():
    txt = open(sys.argv[1]).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    skip = "--skip-peaks" in sys.argv
    agg = collections.OrderedDict()
    for x in rows:
        k = x["Kernel Name"]
        if skip and "_peak_kernel" in k:
            continue
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1e3 if x["Metric Unit"] == "ns" else v * 1e3 if x["Metric Unit"] == "ms" else v
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised under ncu: compare shares)")
    print(f"{'mean us':>10} {'count':>6} {'share':>7}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / n:10.1f} {n:6d} {100 * t / tot:6.1f}%  {k[:110]}")


if __name__ == "__main__":
    main()
