#!/usr/bin/env python
"""Stall reasons and hottest source lines of an ncu source page.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
       python profiles/ncu_stalls.py src.csv [top] [units]     (units = work items of the launch, for "per 32" counts)
"""
import csv
import sys


def main(path, top=22, units=1e8):
    rows = list(csv.reader(open(path, errors="replace")))
    fname, hdr, out = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].isdigit():
            continue
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            out.append((int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0), fname, int(r[0]), r[1].strip(), d))
        except (ValueError, KeyError):
            pass
    tn = sum(o[0] for o in out)
    ts = sum(o[1] for o in out) or 1
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    tot = {k: 0 for k in stalls}
    for o in out:
        for k in stalls:
            try:
                tot[k] += int(o[5].get(k) or 0)
            except ValueError:
                pass
    print(f"warp instructions {tn} ({tn / units * 32:.1f} per 32 units), samples {ts}")
    print({k: round(100 * v / ts, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
    for n, s, f, ln, src, d in sorted(out, key=lambda o: -o[1])[:top]:
        t2 = sorted(((k, int(d.get(k) or 0)) for k in stalls), key=lambda kv: -kv[1])[:2]
        print(f"{100 * s / ts:5.1f}%smp {n / units * 32:6.1f}/32 {f}:{ln}: {src[:80]}  {t2}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22, float(sys.argv[3]) if len(sys.argv) > 3 else 1e8)
