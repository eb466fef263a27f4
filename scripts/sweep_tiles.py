"""Sweep n-tile width and cta_group for the mid-size GEMM shapes (tile-plan overrides through ctrlv_igemm_override) against the heuristic's own choice."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF = torch.bfloat16; dev = "cuda"
torch.manual_seed(0)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=7):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def sweep(name, fn, N):
    ops.lib().ctrlv_igemm_override(0, 0, 0)
    base = t(fn)
    best = (base, "auto")
    line = [f"auto {base:6.1f}"]
    for cg in (1, 2):
        for bn in (64, 96, 128, 160, 192, 256):
            if N % bn and bn != 256: continue
            if bn > N: continue
            ops.lib().ctrlv_igemm_override(bn, cg, 0)
            try:
                v = t(fn)
            except Exception as e:
                v = float("nan")
            line.append(f"cg{cg}/bn{bn} {v:6.1f}")
            if v == v and v < best[0]: best = (v, f"cg{cg}/bn{bn}")
    ops.lib().ctrlv_igemm_override(0, 0, 0)
    print(f"{name:34s} best {best[1]:10s} {best[0]:6.1f} us  ({100*(base-best[0])/base:4.1f}% vs auto) | " + "  ".join(line), flush=True)
def lin(M, K, N, res=False, geglu=False):
    a = torch.randn(M, K, device=dev).to(BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF); b = torch.randn(N, device=dev)
    No = N // 2 if geglu else N
    out = torch.empty(M, No, device=dev, dtype=BF)
    kw = dict(bias=b, out=out, geglu=geglu)
    if res: kw["res1"] = torch.randn(M, No, device=dev).to(BF)
    return lambda: ops.linear(a, w, **kw)
def conv(F_, H, W, C, N):
    x = torch.randn(F_ * H * W, C, device=dev).to(BF); w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).to(BF); b = torch.randn(N, device=dev)
    out = torch.empty(F_ * H * W, N, device=dev, dtype=BF)
    return lambda: ops.conv3x3(x, F_, H, W, w, bias=b, out=out)
def convt(B, T, HW, C, N):
    x = torch.randn(B * T * HW, C, device=dev).to(BF); w = (torch.randn(N, 3 * C, device=dev) / (3 * C) ** 0.5).to(BF); b = torch.randn(N, device=dev)
    out = torch.empty(B * T * HW, N, device=dev, dtype=BF)
    return lambda: ops.conv_t3(x, B, T, HW, w, bias=b, out=out)
sweep("lin 4480x1280->1280 +res (x21)", lin(4480, 1280, 1280, res=True), 1280)
sweep("lin 4480x1280->3840 (x14)", lin(4480, 1280, 3840), 3840)
sweep("lin 17920x640->640 +res (x21)", lin(17920, 640, 640, res=True), 640)
sweep("lin 17920x640->1920 (x14)", lin(17920, 640, 1920), 1920)
sweep("lin 17920x2560->640 +res (x21)", lin(17920, 2560, 640, res=True), 640)
sweep("lin 4480x5120->1280 +res (x21)", lin(4480, 5120, 1280, res=True), 1280)
sweep("geglu 4480x1280->10240 (x21)", lin(4480, 1280, 10240, geglu=True), 10240)
sweep("lin 1120x1280->1280 (x12)", lin(1120, 1280, 1280, res=True), 1280)
sweep("conv 28x5x8 1280->1280 (x16)", conv(28, 5, 8, 1280, 1280), 1280)
sweep("conv 28x10x16 1280->1280 (x5)", conv(28, 10, 16, 1280, 1280), 1280)
sweep("conv 28x20x32 640->640 (x4)", conv(28, 20, 32, 640, 640), 640)
sweep("convt 2x14x2560 320->320 (x14)", convt(2, 14, 2560, 320, 320), 320)
sweep("convt 2x14x640 640->640 (x14)", convt(2, 14, 640, 640, 640), 640)
sweep("convt 2x14x160 1280->1280 (x14)", convt(2, 14, 160, 1280, 1280), 1280)
sweep("convt 2x14x40 1280->1280 (x22)", convt(2, 14, 40, 1280, 1280), 1280)
