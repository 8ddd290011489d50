"""Summarise an `ncu --page raw --csv` export: per kernel launch the metrics that decide what bounds it."""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaH%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor%"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "imma%"),
    ("dram__bytes_read.sum", "dramR"),
    ("dram__bytes_write.sum", "dramW"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "inst"),
    ("launch__occupancy_limit_registers", "occ_regs"),
]


def main(path, substr=None):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx.get("Kernel Name")
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        if substr and substr not in r[name_i]:
            continue
        out = [r[name_i][:44].ljust(44)]
        for k, short in KEYS:
            if k in idx:
                v = r[idx[k]].replace(",", "")
                try:
                    f = float(v)
                    u = units[idx[k]]
                    if short in ("dramR", "dramW"):
                        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                        out.append(f"{short}={f * scale / 1e9:.2f}GB")
                    elif short == "time":
                        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)
                        out.append(f"{f * scale:.3f}ms")
                    elif short == "inst":
                        out.append(f"inst={f:.3g}")
                    else:
                        out.append(f"{short}={f:.1f}")
                except ValueError:
                    pass
        print(" ".join(out))
    if len(sys.argv) > 3:   # list metric names containing a pattern
        for h in hdr:
            if sys.argv[3] in h:
                print(h)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] else None)
