"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} n={v[0]:5d} total={v[1] / 1e6:9.3f} ms share={v[1] / tot * 100:5.1f}% avg={v[1] / v[0] / 1e3:9.1f} us")
print(f"total {tot / 1e6:.3f} ms")
