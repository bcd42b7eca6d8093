"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum, dram__bytes_read.sum,
dram__bytes_write.sum per launch) by kernel: launches, total / average time, share of the step and
average DRAM bytes per launch.  Writes profiles/<out>.json and prints a table.

    python tools/ncu_summary.py gpurun_out/launches.csv profiles/r01_dram_by_kernel.json
"""
import collections
import csv
import json
import re
import sys


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = csv.DictReader(lines)
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    launches = collections.defaultdict(set)
    for r in rows:
        name = r["Kernel Name"]
        m = re.search(r"(\w+_kernel|\w+Kernel\w*)", name)
        key = m.group(1) if m else name[:48]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        metric = r["Metric Name"]
        if metric == "gpu__time_duration.sum":
            v = v / 1e3 if unit in ("nsecond", "ns") else (v * 1e3 if unit in ("msecond", "ms") else v)   # -> us
        elif unit in ("Kbyte", "KB"):
            v *= 1e3
        elif unit in ("Mbyte", "MB"):
            v *= 1e6
        elif unit in ("Gbyte", "GB"):
            v *= 1e9
        per[key][metric] += v
        launches[key].add(r["ID"])
    per.pop("spin_kernel", None)                 # bench.py's head-start kernel of the per-kernel pass: not part of a step
    total = sum(d["gpu__time_duration.sum"] for d in per.values())
    res = {}
    for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(launches[k])
        t = d["gpu__time_duration.sum"]
        dram = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        res[k] = {"launches": n, "total_us": t, "avg_us": t / n, "share": t / total if total else 0.0,
                  "dram_bytes_per_launch": dram / n, "dram_gbs": dram / t / 1e3 if t else 0.0}
        print(f"{k[:44]:44s} n={n:5d} total={t / 1e3:9.3f} ms share={100 * t / total:5.1f}% avg={t / n:8.1f} us "
              f"dram/launch={dram / n / 1e6:8.2f} MB  {dram / t / 1e3 if t else 0:7.0f} GB/s")
    with open(out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
