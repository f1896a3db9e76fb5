#!/usr/bin/env python
"""Turn what profiles/collect_r02.sh left in gpurun_out/ into the committed evidence of round 2:  python profiles/postprocess_r02.py r02"""
import collections
import csv
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
GO = os.path.join(os.path.dirname(HERE), "gpurun_out")


def main(R):
    for f in [f"launches_{R}.csv", f"bench_ref_{R}.json"] + [f"bench_{c}_{R}.json" for c in ("C2", "C3", "C4", "C5")] + \
             [f"ncu_full_{R}_{x}_summary.txt" for x in ("c4", "c2", "c3", "generic")] + [f"ncu_lines_{R}_{x}.txt" for x in ("c4", "c2", "c3", "generic")]:
        shutil.copy(os.path.join(GO, f), os.path.join(HERE, f))
    rows = list(csv.reader(l for l in open(os.path.join(HERE, f"launches_{R}.csv"), errors="replace") if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("qhg::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    bench = json.loads(open(os.path.join(HERE, f"bench_C4_{R}.json")).read().strip().splitlines()[-1])
    km = bench["roofline"]["kernels_ms_per_step"]
    ktot = sum(km.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 (C4: 1e8 agents, 655,362 cells, 1 B200)",
           "cold-cache, serialised launch times: compare SHARES with bench.py's kernels_ms_per_step (last column: the plain run of the same box), not absolutes",
           "k_actions / k_scatter / k_weights_* run once (binning the uploaded agents, first-step weights); k_counts_u64 belongs to the end-to-end loop",
           "", "kernel, launches, total ms, share under ncu, share in the plain bench"]
    alias = {"k_seg_decide<1, 4, 0, 0, 0>": "k_cell_decide", "k_cell_scatter<0, 384, 6, 4, 1, 0>": "k_cell_scatter", "k_cell_scatter<0, 384, 6, 4, 1, 0, 0>": "k_cell_scatter"}
    step_names = set(km) | set(alias)
    stot = sum(ms for n, (c, ms) in agg.items() if alias.get(n, n) in km)
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        b = km.get(alias.get(n, n))
        out.append(f"{n}, {c}, {ms:.3f}, {100 * ms / tot:.1f}% ({100 * ms / stot:.1f}% of the step's kernels), " + (f"{100 * b / ktot:.1f}%" if b else "-"))
    open(os.path.join(HERE, f"launches_{R}_summary.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
