import csv,sys
from collections import defaultdict
lines=[l for l in open(sys.argv[1]) if not l.startswith("==")]
rows=[(r["Kernel Name"].split("(")[0][:55], float(r["Metric Value"].replace(",",""))) for r in csv.DictReader(lines)]
# second half (second cycle)
rows=rows[len(rows)//2:]
tot=defaultdict(lambda:[0,0.0])
for n,v in rows: tot[n][0]+=1; tot[n][1]+=v
al=sum(v for _,v in tot.values())
print("launches %d sum %.1f us"%(len(rows), al/1e3))
for n,(c,v) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:28]: print("%-56s %5d %9.1f %5.1f%%  %6.1f us/launch"%(n,c,v/1e3,100*v/al, v/1e3/c))
