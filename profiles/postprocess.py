#!/usr/bin/env python
"""Turn what profiles/collect.sh left in gpurun_out/ into the committed summaries of a round.
    python profiles/postprocess.py r01
Writes profiles/{launches_R.csv, launches_R_summary.txt, ncu_full_R_summary.txt, traffic_R.json, bench_R.json, clocks_R.csv}."""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main(R):
    go = os.path.join(ROOT, "gpurun_out")
    for f in (f"launches_{R}.csv", f"clocks_{R}.csv", f"bench_{R}.json"):
        shutil.copy(os.path.join(go, f), os.path.join(HERE, f))
    # launch list -> shares
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
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 (1e8 agents, 655,362 cells, 1 B200)",
           "cold-cache, serialised launch times: compare SHARES with bench.py's kernels_ms_per_step, not absolutes",
           "k_actions/k_scatter/k_weights_* run once (binning the uploaded agents, first-step weights); k_counts_u64 belongs to the end-to-end loop",
           "", "kernel, launches, total ms, share"]
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n}, {c}, {ms:.3f}, {100 * ms / tot:.1f}%")
    open(os.path.join(HERE, f"launches_{R}_summary.txt"), "w").write("\n".join(out) + "\n")
    # full capture -> summary + traffic
    rep = os.path.join(go, f"prof_{R}.ncu-rep")
    summ = subprocess.run([sys.executable, os.path.join(HERE, "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    open(os.path.join(HERE, f"ncu_full_{R}_summary.txt"), "w").write(summ)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    traffic = {"round": R, "source": f"profiles/ncu_full_{R}_summary.txt (ncu --set full, bench.py workload, third step)"}
    bench = json.loads(open(os.path.join(HERE, f"bench_{R}.json")).read().strip().splitlines()[-1])
    traffic["agents"] = 0.9925 * bench["config"]["agents_start"]  # third step of the run: the population shrinks ~0.25 % per step
    for r in rr[2:]:
        d = dict(zip(h, r))
        name = "k_cell_decide" if "k_cell_decide" in d["Kernel Name"] else "k_cell_scatter"
        def val(m):
            v, u = float(d[m].replace(",", "")), units[h.index(m)]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        traffic[name] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                         "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1e3, "us": 1.0, "ns": 1e-3}.get(units[h.index("gpu__time_duration.sum")].replace("second", "s").replace("msecond", "ms"), 1.0)}
    json.dump(traffic, open(os.path.join(HERE, f"traffic_{R}.json"), "w"), indent=1)
    print("\n".join(out))
    print(summ)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r01")
