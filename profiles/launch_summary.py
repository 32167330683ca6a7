"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python profiles/launch_summary.py file.csv [n_passes]"""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
npass = int(sys.argv[2]) if len(sys.argv) > 2 else 1
agg = collections.OrderedDict()
for x in rows:
    k = x["Kernel Name"][:64]
    v = float(x["Metric Value"])
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"| `{k}` | {n} | {v / 1e6:.3f} | {v / tot * 100:.1f}% | {v / n / 1e3:.1f} |")
print(f"\ntotal {tot / 1e6:.3f} ms over all captured launches")
