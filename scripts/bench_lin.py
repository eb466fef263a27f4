"""CUDA-event timings of the level-0 short-K linears (the epilogue/L2-bound igemm shapes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF = torch.bfloat16; dev = "cuda"
torch.manual_seed(0)
def run(name, M, K, N, **kw):
    a = torch.randn(M, K, device=dev).to(BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    b = torch.randn(N, device=dev)
    res = kw.pop("res", False)
    if res: kw["res1"] = torch.randn(M, N // (2 if kw.get("geglu") else 1), device=dev).to(BF)
    out = torch.empty(M, N // (2 if kw.get("geglu") else 1), device=dev, dtype=BF)
    big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3): ops.linear(a, w, bias=b, out=out, **kw)
    ts = []
    for _ in range(10):
        big.zero_()  # flush L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(a, w, bias=b, out=out, **kw); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    hot = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(a, w, bias=b, out=out, **kw); e1.record(); torch.cuda.synchronize()
        hot.append(e0.elapsed_time(e1) * 1e3)
    fl = 2.0 * M * K * N
    print(f"{name:28s} cold {sorted(ts)[len(ts)//2]:7.1f} us  hot {sorted(hot)[len(hot)//2]:7.1f} us  {fl/sorted(hot)[len(hot)//2]/1e6:7.1f} TF/s hot")
run("res320 71680x320->320 +res", 71680, 320, 320, res=True)
run("lin320 71680x320->320", 71680, 320, 320)
run("qkv 71680x320->960", 71680, 320, 960)
run("geglu 71680x320->2560", 71680, 320, 2560, geglu=True)
run("ffdown 71680x1280->320 +res", 71680, 1280, 320, res=True)
run("res640 17920x640->640 +res", 17920, 640, 640, res=True)
run("geglu 17920x640->5120", 17920, 640, 5120, geglu=True)
