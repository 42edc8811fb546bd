"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
   python tools/summarize_launches.py launches.csv [skip_first_n_launches] > profiles/<name>.md"""
import csv, collections, re, sys
path = sys.argv[1]; skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.DictReader(lines))[skip:]
agg = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"].replace("void ", "").replace("stg::<unnamed>::", "").replace("stg::", "")
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    a = agg.setdefault(name, [0, 0.0, 0.0])
    t = float(r["Metric Value"].replace(",", "")) / 1e3
    a[0] += 1; a[1] += t; a[2] = max(a[2], t)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total {tot/1e3:.3f} ms (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total us | share | avg us | max us |\n|---|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {a[0]} | {a[1]:.1f} | {100*a[1]/tot:.1f}% | {a[1]/a[0]:.1f} | {a[2]:.1f} |")
