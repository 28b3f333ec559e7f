"""Largest individual launches of an ncu gpu__time_duration launch list (second half = steady cycle)."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = [(int(r["ID"]), r["Kernel Name"].split("(")[0][:44], float(r["Metric Value"].replace(",", "")) / 1e3,
         r["Grid Size"], r["Block Size"]) for r in csv.DictReader(lines)]
rows = rows[len(rows) // 2:]
print("launches %d  sum %.1f us" % (len(rows), sum(r[2] for r in rows)))
for r in sorted(rows, key=lambda r: -r[2])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%6d %-45s %9.1f us  grid %s block %s" % r)
