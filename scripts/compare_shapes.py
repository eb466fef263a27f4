import re, sys
def load(f):
    d = {}
    for l in open(f):
        m = re.match(r"\s*([\d.]+) ms n=\s*(\d+) avg=\s*([\d.]+) us\s+([\d.]+) TF\s+(\w+) (\(.*\))", l)
        if m: d[(m.group(5), m.group(6))] = (float(m.group(1)), int(m.group(2)), float(m.group(3)), float(m.group(4)))
    return d
files = sys.argv[1:]
ds = [load(f) for f in files]
keys = [k for k in ds[0] if k[0] in ("linear", "conv3x3", "conv_t3") and all(k in d for d in ds)]
tot = [sum(d[k][0] for k in keys) for d in ds]
best = sum(min(d[k][0] for d in ds) for k in keys)
print("totals:", [f"{t:.2f}" for t in tot], "best-of:", f"{best:.2f}")
for k in sorted(keys, key=lambda k: -ds[0][k][0])[:int(40)]:
    print(" | ".join(f"{d[k][2]:7.1f}us {d[k][3]:5.0f}TF" for d in ds), f" n={ds[0][k][1]:3d} {k[0]} {k[1]}")
