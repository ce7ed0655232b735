"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=40):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    tot = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns, _, _ in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    print("launches %d, total %.3f ms" % (len(rows), tot / 1e6))
    print("%-60s %6s %10s %7s" % ("kernel", "count", "ms", "share"))
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-60s %6d %10.3f %6.1f%%" % (n[:60], c, ns / 1e6, 100 * ns / tot))
    return rows


if __name__ == "__main__":
    rows = main(sys.argv[1])
    if len(sys.argv) > 2:
        pat = sys.argv[2]
        for i, (n, ns, g, b) in enumerate(rows):
            if re.search(pat, n):
                print(i, n[:50], g, b, "%.1f us" % (ns / 1e3))
