"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / mean / share."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    name = row['Kernel Name'].split('(')[0]
    agg[(name, row['Grid Size'], row['Block Size'])].append(float(row['Metric Value'].replace(',', '')) / 1e3)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':70s} {'grid':>14s} {'block':>12s} {'n':>4s} {'mean us':>9s} {'share':>6s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k[0][:70]:70s} {k[1]:>14s} {k[2]:>12s} {len(v):4d} {sum(v)/len(v):9.2f} {100*sum(v)/tot:5.1f}%")
