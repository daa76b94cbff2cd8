#!/usr/bin/env python3
"""Summarise the per-instruction stall samples of one kernel from `ncu --page source --csv`:
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > k.csv ; python tools/ncu_stalls.py k.csv [topN]"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    data = []
    for r in rows[hi + 1:]:                      # first kernel instance only
        if r and r[0] in ("Kernel Name", "Address"):
            break
        if len(r) == len(hdr):
            data.append(r)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    si = hdr.index("# Samples"); ie = hdr.index("Instructions Executed")
    tot = sum(int(r[si]) for r in data)
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
    print("total samples", tot, "kernel", rows[0][1][:80])
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    ops = {}
    for r in data:
        op = r[1].split()[0] if not r[1].strip().startswith("@") else r[1].split()[1]
        ops[op] = ops.get(op, 0) + int(r[ie] or 0)
    ti = sum(ops.values())
    print("instruction mix:", {k: f"{100 * v / ti:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]})
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -int(r[si]))[:top]:
        st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i] or 0)}
        print(f"{int(r[si]):6d} {r[1].strip()[:70]:70s} {dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])}")


if __name__ == "__main__":
    main()
