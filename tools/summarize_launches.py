#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table."""
import collections
import csv
import io
import sys


def main(path, out):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        agg.setdefault(row["Kernel Name"], []).append(v)
        n += 1
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"| total ms | launches | avg ms | share | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {sum(v):.3f} | {len(v)} | {sum(v)/len(v):.4f} | {100*sum(v)/tot:.1f}% | `{k[:110]}` |\n")
        f.write(f"\n{n} launches, {tot:.3f} ms of kernel time (cold-cache, serialised under ncu: compare shares, not absolutes)\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
