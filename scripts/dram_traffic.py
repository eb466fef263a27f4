"""ncu launch list of one eager step (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; see
scripts/gpu_evidence.sh) -> profiles/r02_step_dram_traffic.json: per-kernel launches, time and DRAM bytes, and the
mean DRAM traffic per launch of the tensor-core contraction kernels (bench.py's roofline.traffic).
   python scripts/dram_traffic.py gpurun_out/step_launches_dram.csv profiles/r02_step_dram_traffic.json"""
import collections
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
        "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def main(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    k = collections.defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
    for r in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("ctrlv::", "").strip()
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            k[name]["launches"] += 1
            k[name]["ms"] += v
        elif m == "dram__bytes_read.sum":
            k[name]["dram_read_bytes"] += v
        elif m == "dram__bytes_write.sum":
            k[name]["dram_write_bytes"] += v
    tensor = [n for n in k if "igemm_kernel" in n or "ff_kernel" in n]
    n_t = sum(k[n]["launches"] for n in tensor)
    b_t = sum(k[n]["dram_read_bytes"] + k[n]["dram_write_bytes"] for n in tensor)
    out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over "
                     "one eager step (scripts/profile_step.py), 14x320x512 CFG batch 2, round 2 (GroupNorm statistics in the "
                     "producers' epilogues, fused FeedForward at level 0)",
           "launches": sum(v["launches"] for v in k.values()),
           "ms_serialised": sum(v["ms"] for v in k.values()),
           "dram_bytes_per_step": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for v in k.values()),
           "igemm_launches": n_t, "igemm_dram_bytes_per_step": b_t,
           "igemm_dram_bytes_per_launch": b_t / max(n_t, 1),
           "kernels": dict(sorted(k.items(), key=lambda kv: -kv[1]["ms"]))}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({a: out[a] for a in ("launches", "ms_serialised", "dram_bytes_per_step", "igemm_launches",
                                          "igemm_dram_bytes_per_launch")}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
