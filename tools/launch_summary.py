"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (shares of the total)."""
import collections, csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    tot[k] += v; cnt[k] += 1
T = sum(tot.values())
print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{k[:44]:44s} {cnt[k]:8d} {v/1e3:10.3f} {v/cnt[k]:9.1f} {v/T*100:6.1f}%")
