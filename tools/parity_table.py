#!/usr/bin/env python
"""profiles/r02_force_parity.md from the report `pytest -m gpu` writes (tests/conftest.py::record_parity):
    python tools/parity_table.py profiles/r02_force_parity_report.json > profiles/r02_force_parity.md"""
import json
import sys


def main(path: str) -> int:
    recs = json.load(open(path))
    print("# Round 2: measured parity margins of every GPU force / energy comparison\n")
    print("Written by `pytest -m gpu` (tests/conftest.py::record_parity -> gpurun_out/force_parity_report.json) on one B200; this is the\n"
          "committed copy of the final run of the round (87 GPU tests green; `python tools/parity_table.py` makes this table).  Units:\n"
          "multiples of north_star's force tolerance `1e-5 |F| + 1e-6 eV/A` at the WORST element; `ref32` = the reference-precision\n"
          "result (the oracle in fp32 / the fixture written by the reference's own source), `fp64` = the all-fp64 oracle.  rms errors\n"
          "in eV/A against fp64.  Criterion `literal`: worst element <= 1.0 against ref32 AND fp64; `relative`: see tests/conftest.py.\n")
    print("| comparison | elements | kernel vs ref32 | kernel vs fp64 | ref32 vs fp64 | rms err kernel | rms err ref32 | criterion |")
    print("|---|---:|---:|---:|---:|---:|---:|---|")
    others = []
    for where, r in sorted(recs.items()):
        if "kernel_vs_ref32_x_tol" not in r:
            others.append((where, r))
            continue
        f = lambda k, fmt: (fmt % r[k]) if k in r else "—"
        print(f"| {where} | {int(r.get('n_elements', 0))} | {f('kernel_vs_ref32_x_tol', '%.3f')} | {f('kernel_vs_fp64_x_tol', '%.3f')} | "
              f"{f('ref32_vs_fp64_x_tol', '%.3f')} | {f('kernel_rms_err', '%.2e')} | {f('ref32_rms_err', '%.2e')} | {r.get('criterion', '')} |")
    if others:
        print("\nOther recorded comparisons (energies, relative errors):\n")
        for where, r in others:
            print(f"* {where}: " + ", ".join(f"{k} = {v:.3e}" if isinstance(v, float) else f"{k} = {v}" for k, v in sorted(r.items())))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
