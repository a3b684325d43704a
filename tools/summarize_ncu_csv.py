"""Turns an ncu `--metrics ... --csv --log-file F` capture (one row per kernel launch and metric) into a markdown table,
one row per launch.   python tools/summarize_ncu_csv.py F "title" > profiles/rNN_x.md"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "ms", 1e-6), ("smsp__inst_executed.sum", "warp instr (M)", 1e-6),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1),
        ("dram__bytes_read.sum", "DRAM rd MB", None), ("dram__bytes_write.sum", "DRAM wr MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("lts__t_bytes.sum", "L2 MB", None),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %", 1), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %", 1),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %", 1), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 %", 1),
        ("launch__registers_per_thread", "regs", 1)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def load(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    rows = {}
    order = []
    for r in csv.DictReader(lines):
        k = int(r["ID"])
        if k not in rows:
            name = re.sub(r"^(void )?<unnamed>::", "", r["Kernel Name"])
            rows[k] = {"name": re.sub(r"\(.*$", "", name), "grid": r["Grid Size"], "block": r["Block Size"]}
            order.append(k)
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"].strip()
        name = r["Metric Name"]
        if "bytes" in name:
            v *= UNIT.get(u, 1e-6)                 # -> MB
        elif name == "gpu__time_duration.sum":
            v *= UNIT.get(u, 1e-6)                 # -> ms
        elif name == "smsp__inst_executed.sum":
            v *= 1e-6
        rows[k][name] = v
    return [rows[k] for k in order]


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "ncu counters")
    rows = load(path)
    print("# " + title + "\n")
    print("| # | kernel | grid | block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|---|---|" + "---|" * len(COLS))
    for i, r in enumerate(rows):
        cells = []
        for m, _, _ in COLS:
            v = r.get(m)
            cells.append("-" if v is None else ("%.3f" % v if m == "gpu__time_duration.sum" else ("%.0f" % v if v >= 100 else "%.1f" % v)))
        print("| %d | `%s` | %s | %s | %s |" % (i, r["name"], r["grid"], r["block"], " | ".join(cells)))
    tot = sum(r.get("gpu__time_duration.sum", 0) for r in rows)
    ins = sum(r.get("smsp__inst_executed.sum", 0) for r in rows)
    print("\nTotal: %.2f ms of kernel time, %.0f M warp instructions." % (tot, ins))


if __name__ == "__main__":
    main()
