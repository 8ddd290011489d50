#!/usr/bin/env python
"""Top stall-sample instructions of a kernel from an `ncu --set full --import-source on` report.

    python tools/ncu_source_hotspots.py gpurun_out/r01k_pair.ncu-rep k_pair_fwd [--top 20] [--context 12]

Reads `ncu -i REPORT --page source --csv --kernel-name regex:KERNEL` and prints, per captured launch, the SASS instructions
that collected the most warp-stall samples, with their share of the kernel's samples and their execution counts.  An
instruction far above the rest is where the warps wait: look at what it reads (--context prints the instructions in front
of the hottest one).  This is how the write-after-write on the neighbour gather in k_pair_fwd (28 % of the samples on one
SEL) and the exposed candidate load of k_nl_scan (38 %) were found (profiles/r01k_ncu_pair_summary.md).
"""
import argparse
import csv
import subprocess
import sys


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel", help="regex for --kernel-name")
    ap.add_argument("--top", type=int, default=20)
    ap.add_argument("--context", type=int, default=0, help="also print this many instructions before the hottest one")
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--kernel-name", f"regex:{args.kernel}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for b in blocks:
        if not b["rows"] or b["name"] in seen:      # the page repeats each kernel for its second view
            continue
        seen.add(b["name"])
        hdr, body = b["rows"][0], b["rows"][1:]
        i_src, i_all = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
        i_exec, i_long = hdr.index("Instructions Executed"), hdr.index("stall_long_sb")
        body = body[: len(body) // 2] if len(body) % 2 == 0 and body[0][i_src] == body[len(body) // 2][i_src] else body
        data = [(int(r[i_all]), idx, r) for idx, r in enumerate(body) if len(r) > i_all and r[i_all].isdigit()]
        total = sum(d[0] for d in data) or 1
        print(f"== {b['name'][:100]}\n   {total} stall samples over {len(data)} instructions")
        ranked = sorted(data, key=lambda d: -d[0])[: args.top]
        for n, idx, r in ranked:
            print(f"{n:8d} {100.0 * n / total:5.1f}%  long_sb={r[i_long]:>6}  #{idx:5d} exec={r[i_exec]:>10}  {r[i_src].strip()[:100]}")
        if args.context and ranked:
            hot = ranked[0][1]
            print(f"   -- {args.context} instructions in front of #{hot}:")
            for idx in range(max(0, hot - args.context), hot + 1):
                r = body[idx]
                print(f"{r[i_all]:>8}         #{idx:5d} exec={r[i_exec]:>10}  {r[i_src].strip()[:100]}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
