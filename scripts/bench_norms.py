"""CUDA-event timing of the HBM-bound kernels (GroupNorm apply, LayerNorm, residual add + statistics, temporal attention)
at the row counts of BASELINE configs 2 / 3 / 4 (level 0 and 1), 256 MB L2 flush between calls: GB/s on the algorithmic
byte count (read once + write once) against the measured HBM copy peak."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF, dev = torch.bfloat16, "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=8):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


CASES = [("c2 L0", 28, 2560, 320), ("c2 L1", 28, 640, 640), ("c3 L0", 224, 2560, 320), ("c4 L0", 50, 9216, 320),
         ("c4 L1", 50, 2304, 640), ("c4 L2", 50, 576, 1280)]
only = os.environ.get("CASES")
for name, units, rows, C in CASES:
    if only and not any(o in name for o in only.split(",")): continue
    M = units * rows
    x = torch.randn(M, C, device=dev).to(BF); o = torch.empty_like(x)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    st = ops.GNStats(torch.zeros(ops.GNStats.numel(units), dtype=torch.int64, device=dev), units, rows, C)
    ops.gn_stats_of(x, (st, 0))
    nbytes = 4.0 * M * C
    r = {"case": name, "rows": M, "C": C, "MB": round(nbytes / 1e6, 1)}
    us = timeit(lambda: ops.groupnorm(x, units, rows, g, b, 1e-5, True, out=o, stats=st)); r["gn_apply_us"] = round(us, 1); r["gn_apply_gbs"] = round(nbytes / us / 1e3)
    us = timeit(lambda: ops.layernorm(x, out=o)); r["ln_us"] = round(us, 1); r["ln_gbs"] = round(nbytes / us / 1e3)
    us = timeit(lambda: ops.axpby(x, x, out=o)); r["axpby_us"] = round(us, 1); r["axpby_gbs"] = round(6.0 * M * C / us / 1e3)
    us = timeit(lambda: o.copy_(x)); r["torch_copy_us"] = round(us, 1); r["torch_copy_gbs"] = round(nbytes / us / 1e3)
    print(json.dumps(r), flush=True)
    del x, o
    torch.cuda.empty_cache()
