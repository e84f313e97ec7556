"""Developer tool: `ncu --metrics gpu__time_duration.sum --csv ... | python tools/ncu_times.py` -> kernel, us per launch."""
import csv, sys
rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
h = rows[0]
for r in rows[1:]:
    print(f'{r[h.index("Kernel Name")].split("(")[0][:60]:60s} {r[h.index("Metric Name")]:28s} {float(r[h.index("Metric Value")].replace(",", "")) / 1e3:10.2f} us')
