"""Summarise an ncu `--csv` launch list (gpu__time_duration.sum) by kernel name."""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if unit in ("ns", "nsecond"): v /= 1e3
    elif unit in ("ms", "msecond"): v *= 1e3
    elif unit in ("s", "second"): v *= 1e6
    tot[name][0] += 1; tot[name][1] += v
s = sum(v[1] for v in tot.values())
print(f"total {s/1e3:.2f} ms over {sum(v[0] for v in tot.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:9.3f} ms  {100*v[1]/s:5.1f}%  n={v[0]:5d}  avg={v[1]/v[0]:8.1f} us  {k}")
