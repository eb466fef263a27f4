"""CUDA-event timing of the spatial / temporal attention cores at the step's shapes (256 MB L2 flush between calls)."""
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


for (F_, S, heads) in ((28, 2560, 5), (28, 640, 10), (28, 160, 20), (28, 40, 20)):
    C = heads * 64
    qkv = torch.randn(F_ * S, 3 * C, device=dev).to(BF)
    out = torch.empty(F_ * S, C, device=dev, dtype=BF)
    us = timeit(lambda: ops.attn_spatial(qkv, F_, S, heads, out=out))
    fl = 4.0 * F_ * heads * S * S * 64
    print(json.dumps({"case": f"spatial S={S} heads={heads}", "us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}), flush=True)
    B, T = 2, F_ // 2
    us = timeit(lambda: ops.attn_temporal(qkv, B, T, S, heads, out=out))
    print(json.dumps({"case": f"temporal S={S} heads={heads} T={T}", "us": round(us, 1),
                      "gbs": round(8.0 * F_ * S * C / us / 1e3, 1)}), flush=True)
