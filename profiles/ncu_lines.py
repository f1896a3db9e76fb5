#!/usr/bin/env python
"""Per-CUDA-line summary of an ncu source page.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > src.csv
       python profiles/ncu_lines.py src.csv [top]
"""
import csv
import sys


def main(path, top=30):
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
            out.append((int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), fname, int(r[0]), r[1].strip()))
        except (ValueError, KeyError):
            pass
    ts = sum(o[0] for o in out) or 1
    tn = sum(o[1] for o in out) or 1
    print(f"total samples {ts}, warp instructions {tn}")
    for s, n, f, ln, src in sorted(out, key=lambda o: -o[0])[:top]:
        print(f"{100*s/ts:5.1f}% smp {100*n/tn:5.1f}% inst  {f}:{ln}: {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
