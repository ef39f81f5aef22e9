"""Per-kernel table of an ncu --csv launch list that carries gpu__time_duration.sum and dram__bytes_{read,write}.sum:
python scripts/ncu_launch_table.py launches.csv"""
import csv, sys, collections, re
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
hdr = lines[0]
r = csv.DictReader([hdr] + [l for l in lines[1:] if l != hdr])
per = collections.OrderedDict()
for row in r:
    e = per.setdefault(row["ID"], {"k": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", ""), "t": 0.0, "b": 0.0})
    v = float(row["Metric Value"].replace(",", ""))
    if "time" in row["Metric Name"]:
        u = row["Metric Unit"]
        e["t"] = v / 1000.0 if u.startswith("n") else (v if u.startswith("u") else v * 1000.0)
    else:
        e["b"] += v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for e in per.values():
    a = agg[e["k"]]; a[0] += 1; a[1] += e["t"]; a[2] += e["b"]
tot = sum(a[1] for a in agg.values()); totb = sum(a[2] for a in agg.values())
print(f"total {tot/1000:.2f} ms over {len(per)} launches, DRAM traffic {totb/1e9:.1f} GB")
for k, (n, us, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us/1000:8.3f} ms {100*us/tot:5.1f}%  n={n:4d}  avg {us/n:7.1f} us  dram {b/1e9:6.2f} GB  {b/us/1e6 if us else 0:5.2f} TB/s  {k[:70]}")
