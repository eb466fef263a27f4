"""A few isolated kernel launches for `ncu --set full` captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF = torch.bfloat16; dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "geglu"
torch.manual_seed(0)
def lin(M, K, N, **kw):
    a = torch.randn(M, K, device=dev).to(BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    b = torch.randn(N, device=dev)
    if kw.pop("res", False):
        kw["res1"] = torch.randn(M, N, device=dev).to(BF)
    for _ in range(3): ops.linear(a, w, bias=b, **kw)
    torch.cuda.synchronize()
if which == "geglu": lin(71680, 320, 2560, geglu=True)
elif which == "res320": lin(71680, 320, 320, res=True)
elif which == "ffdown": lin(71680, 1280, 320, res=True)
elif which == "attn":
    qkv = torch.randn(28 * 2560, 960, device=dev).to(BF)
    for _ in range(3): ops.attn_spatial(qkv, 28, 2560, 5)
    torch.cuda.synchronize()
if which in ("ff", "ff1"):
    from ctrlv_b200 import _lib
    _lib.load().ctrlv_feedforward_override(1 if which == "ff1" else 0)
    M, C = 71680, 320
    x = torch.randn(M, C, device=dev).to(BF)
    w1 = (torch.randn(8 * C, C, device=dev) / C ** 0.5).to(BF); b1 = torch.randn(8 * C, device=dev)
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(BF); b2 = torch.randn(C, device=dev)
    res = torch.randn(M, C, device=dev).to(BF)
    for _ in range(3): ops.feedforward(x, w1, b1, w2, bias=b2, res1=res)
    torch.cuda.synchronize()
print("ok")
if which == "norms":
    x = torch.randn(71680, 320, device=dev).to(BF); g = torch.randn(320, device=dev); b = torch.randn(320, device=dev); o = torch.empty_like(x)
    for _ in range(2):
        st = ops.GNStats(torch.zeros(ops.GNStats.numel(28), dtype=torch.int64, device=dev), 28, 2560, 320)
        ops.gn_stats_of(x, (st, 0))
        ops.groupnorm(x, 28, 2560, g, b, 1e-5, True, out=o, stats=st)   # apply pass only (statistics from the producer)
        ops.groupnorm(x, 2, 35840, g, b, 1e-5, True, out=o)            # two-pass form (kept for inputs without a producer)
        ops.layernorm(x, g, b, out=o)
    torch.cuda.synchronize()
if which == "conv":
    x = torch.randn(28 * 40 * 64, 320, device=dev).to(BF); w = (torch.randn(320, 9 * 320, device=dev) / 54).to(BF); b = torch.randn(320, device=dev)
    for _ in range(3): ops.conv3x3(x, 28, 40, 64, w, bias=b)
    x = torch.randn(28 * 20 * 32, 640, device=dev).to(BF); w = (torch.randn(640, 9 * 640, device=dev) / 76).to(BF); b = torch.randn(640, device=dev)
    for _ in range(3): ops.conv3x3(x, 28, 20, 32, w, bias=b)
    torch.cuda.synchronize()
if which == "tattn":
    qkv = torch.randn(28 * 2560, 960, device=dev).to(BF)
    for _ in range(3): ops.attn_temporal(qkv, 2, 14, 2560, 5)
    torch.cuda.synchronize()
if which == "sk":  # stream-K schedule on the level-3 problems (10 row tiles)
    a = torch.randn(1120, 1280, device=dev).to(BF); w = (torch.randn(1280, 1280, device=dev) / 36).to(BF); b = torch.randn(1280, device=dev)
    r = torch.randn(1120, 1280, device=dev).to(BF)
    for _ in range(3): ops.linear(a, w, bias=b, res1=r)
    x = torch.randn(28 * 5 * 8, 1280, device=dev).to(BF); w = (torch.randn(1280, 9 * 1280, device=dev) / 107).to(BF)
    for _ in range(3): ops.conv3x3(x, 28, 5, 8, w, bias=b)
    torch.cuda.synchronize()
if which == "lnqkv":  # LayerNorm + q|k|v projection in one launch (ctrlv_linear_ln), level-0 shape
    x = torch.randn(71680, 320, device=dev).to(BF); w = (torch.randn(960, 320, device=dev) / 18).to(BF); b = torch.randn(960, device=dev)
    o = torch.empty(71680, 960, device=dev, dtype=BF)
    for _ in range(3): ops.linear_ln(x, w, bias=b, out=o)
    torch.cuda.synchronize()
