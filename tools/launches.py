"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel
totals and the launches of the last full generator pass."""
import csv
import sys
from collections import defaultdict


def main(path, per_pass=True):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", "")),
                     r["Grid Size"]))
    tot = defaultdict(lambda: [0, 0.0])
    for _, n, v, _ in rows:
        tot[n][0] += 1
        tot[n][1] += v
    all_ns = sum(v for _, v in tot.values())
    print("%-45s %6s %12s %7s" % ("kernel", "count", "total_us", "share"))
    for n, (c, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("%-45s %6d %12.1f %6.1f%%" % (n[:45], c, v / 1e3, 100 * v / all_ns))
    if per_pass:
        idx = [i for i, r in enumerate(rows) if "pack_ncl" in r[1]]
        if len(idx) >= 2:
            s, e = idx[-2], idx[-1]
            print("\nlast full pass (launch, kernel, us, grid):")
            t = 0
            for r in rows[s:e]:
                print("  %4d %-40s %9.1f %s" % (r[0], r[1][:40], r[2] / 1e3, r[3]))
                t += r[2]
            print("  pass total %.1f us" % (t / 1e3))


if __name__ == "__main__":
    main(sys.argv[1])
