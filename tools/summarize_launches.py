"""Turns an ncu launch list (--metrics gpu__time_duration.sum --csv --log-file F) into a per-kernel share table (markdown).

  python tools/summarize_launches.py gpurun_out/launches.csv "title line" > profiles/rNN_launch_list.md
"""
import csv
import re
import sys


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "kernel launch list")
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^(void )?<unnamed>::", "", r["Kernel Name"])
        name = re.sub(r"\(.*$", "", name)
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r["Metric Unit"].strip(), 1e-6)
        rows.append((name, float(r["Metric Value"].replace(",", "")) * scale, r["Grid Size"], r["Block Size"]))
    total = sum(ms for _, ms, _, _ in rows)
    agg = {}
    for name, ms, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ms
    print("# " + title + "\n")
    print("%d launches, %.1f ms of kernel time in total (cold caches, serialised by the profiler: compare SHARES, not "
          "absolutes).\n" % (len(rows), total))
    print("| kernel | launches | total ms | ms per launch | share | grid (first) | block |")
    print("|---|---|---|---|---|---|---|")
    for name, (n, ms, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.3f | %.1f %% | %s | %s |" % (name, n, ms, ms / n, 100 * ms / total, grid, block))


if __name__ == "__main__":
    main()
