#!/usr/bin/env python
"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X.csv`) -> per-kernel table of the LAST step captured.
    python tools/launch_table.py gpurun_out/r02w_launches.csv > profiles/r02w_launches_table.md
Per-launch times under ncu are cold-cache and serialised: what must agree with bench.py's live per-kernel times is each
kernel's SHARE of the step, not the absolute numbers."""
import csv
import re
import sys
from collections import OrderedDict


def main(path: str) -> int:
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            name = re.sub(r"^void ", "", r["Kernel Name"])
            name = re.sub(r"\(.*$", "", name)
            name = name.replace("apax::<unnamed>::", "").replace("<unnamed>::", "")
            unit = r["Metric Unit"]
            v = float(r["Metric Value"].replace(",", ""))
            ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
            rows.append((name, ms))
    # the last E+F step = everything from the last k_atom_pack on
    starts = [i for i, (n, _) in enumerate(rows) if n.startswith("k_atom_pack")]
    step = rows[starts[-1]:] if starts else rows
    step = [(n, t) for n, t in step if n.startswith("k_")]
    agg = OrderedDict()
    for n, t in step:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    total = sum(v[1] for v in agg.values())
    print(f"Launch list of one energy+forces step under ncu ({path}; {len(step)} launches, {total:.3f} ms serialised, cold cache)\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {t:.3f} | {100 * t / total:.1f} % |")
    print(f"\nAll captured launches: {len(rows)} (list build, warm-up and timed steps).")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
