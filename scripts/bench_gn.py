"""Microbenchmark of GroupNorm statistics placement: for the step's GroupNorm inputs,
time the pair [producer launch -> GroupNorm] with the statistics taken (a) by a gn_stats_kernel pass
(two-pass) and (b) from the producer's epilogue (fused).  CUDA events around the pair, 256 MB L2 flush before
each pair (the norm itself then reads what its producer left in L2, as in the step).
   python scripts/bench_gn.py  -> gpurun_out/gn_pairs.json"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF, dev = torch.bfloat16, "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def case(kind, B, T, H, W, C, temporal):
    """producer of a [B*T*H*W, C] tensor + the GroupNorm(32, C) that consumes it"""
    F_, HW = B * T, H * W
    M = F_ * HW
    units, rows = (B, T * HW) if temporal else (F_, HW)
    g = torch.randn(C, device=dev); be = torch.randn(C, device=dev); bias = torch.randn(C, device=dev)
    out = torch.empty(M, C, device=dev, dtype=BF); nout = torch.empty(M, C, device=dev, dtype=BF)
    x = torch.randn(M, C, device=dev).to(BF)
    res = torch.randn(M, C, device=dev).to(BF)
    if kind == "conv3x3":
        w = (torch.randn(C, 9 * C, device=dev) / (9 * C) ** 0.5).to(BF)
        prod = lambda gn: ops.conv3x3(x, F_, H, W, w, bias=bias, out=out, res1=res if temporal else None, gn=gn)
    elif kind == "conv_t3":
        w = (torch.randn(C, 3 * C, device=dev) / (3 * C) ** 0.5).to(BF)
        prod = lambda gn: ops.conv_t3(x, B, T, HW, w, bias=bias, out=out, res1=None if temporal else res, s_acc=0.5, gn=gn)
    else:
        w = (torch.randn(C, C, device=dev) / C ** 0.5).to(BF)
        prod = lambda gn: ops.linear(x, w, bias=bias, out=out, res1=res, gn=gn)
    st = ops.GNStats(torch.zeros(ops.GNStats.numel(units), dtype=torch.int64, device=dev), units, rows, C)
    t_prod = timeit(lambda: prod(None))
    t_prod_gn = timeit(lambda: prod((st, 0)))
    t_two = timeit(lambda: (prod(None), ops.groupnorm(out, units, rows, g, be, 1e-6, True, out=nout)))
    t_fused = timeit(lambda: (prod((st, 0)), ops.groupnorm(out, units, rows, g, be, 1e-6, True, out=nout, stats=st)))
    return dict(producer_us=round(t_prod, 1), producer_gn_us=round(t_prod_gn, 1), pair_two_pass_us=round(t_two, 1),
                pair_fused_us=round(t_fused, 1), rep=st.rep)


LEVELS = [(0, 40, 64, 320), (1, 20, 32, 640), (2, 10, 16, 1280), (3, 5, 8, 1280)]
# per ResBlock: conv1 -> norm2 (spatial units), conv2(+res) -> tnorm1 (temporal), tconv1 -> tnorm2 (temporal),
# tconv2(+res) -> next norm (spatial); per transformer: proj_out(+res) -> next norm (spatial)
KINDS = [("conv3x3", False), ("conv3x3", True), ("conv_t3", True), ("conv_t3", False), ("linear", False)]
rows = []
only = os.environ.get("LEVELS")
for lvl, H, W, C in LEVELS:
    if only and str(lvl) not in only.split(","): continue
    for kind, temporal in KINDS:
        r = dict(level=lvl, kind=kind, temporal_units=temporal, C=C, **case(kind, 2, 14, H, W, C, temporal))
        rows.append(r)
        print(json.dumps(r), flush=True)
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gn_pairs.json", "w"), indent=1)
