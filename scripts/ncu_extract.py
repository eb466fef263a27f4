"""Compact per-kernel metric extracts from `ncu --set full` reports (read on the CPU with `ncu -i`):
    python scripts/ncu_extract.py gpurun_out/ncu6_conv.ncu-rep ... -> profiles/r02_ncu_raw_<name>.csv
One row per (kernel launch, metric): the columns the roofline statements in profiles/r02_ncu_summary.md
are read from (duration, tensor / XU / FMA pipe activity, issue slots, DRAM and L2 bytes, registers,
occupancy), so the summary can be checked without the multi-MB .ncu-rep files."""
import csv
import io
import os
import re
import subprocess
import sys

KEEP = [r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.max$", r"sm__pipe_tensor.*cycles_active.*pct", r"sm__inst_executed_pipe_(xu|fma|fmaheavy|alu|uniform|lsu)\.?.*pct",
        r"^sm__inst_executed_pipe_tensor", r"^smsp__issue_active\.avg\.pct", r"^sm__warps_active\.avg\.pct_of_peak",
        r"^dram__bytes_(read|write)\.sum$", r"^dram__throughput\.avg\.pct", r"^lts__t_bytes\.sum$", r"^lts__t_sectors_srcunit_tex_op_(read|write)\.sum$",
        r"^lts__throughput\.avg\.pct", r"^l1tex__throughput\.avg\.pct", r"^launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|cluster_size|waves_per_multiprocessor|occupancy_limit_\w+)$",
        r"^sm__throughput\.avg\.pct", r"^smsp__cycles_active\.avg$", r"^sm__ctas_launched\.sum$", r"^smsp__inst_executed\.sum$"]


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        names, units, data = rows[0], rows[1], rows[2:]
        cols = [i for i, n in enumerate(names) if any(re.search(p, n) for p in KEEP)
                and not re.search(r"\.(max|min|sum)\.pct|TriageCompute", n)]
        kcol = names.index("Kernel Name")
        name = re.sub(r"^ncu\d*_|\.ncu-rep$", "", os.path.basename(rep))
        out = os.path.join(root, "profiles", f"r02_ncu_raw_{name}.csv")
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["launch", "kernel", "metric", "unit", "value"])
            for r in data:
                k = re.sub(r"\(.*", "", r[kcol])[:60]
                for i in cols:
                    w.writerow([r[0], k, names[i], units[i], r[i]])
        print(out, len(data), "launches", len(cols), "metrics")


if __name__ == "__main__":
    main()
