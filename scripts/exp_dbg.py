import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF = torch.bfloat16
def bench(M, K, N, iters=10):
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF); out = torch.empty(M, N, device="cuda", dtype=BF)
    for _ in range(3): ops.linear(a, w, out=out)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.linear(a, w, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tiles = ((M + 127) // 128) * (N // int(os.environ.get("CTRLV_DEBUG_BN", 256)))
    kb = K // 64
    cyc = ms * 1e-3 * 1.75e9 / (tiles / 148) / kb
    print(f"DBG={os.environ.get('CTRLV_DEBUG_DBG')} BN={os.environ.get('CTRLV_DEBUG_BN')} M={M} K={K} N={N}: {ms*1e3:.1f} us  ~{cyc:.0f} cycles per k-block per CTA", flush=True)
bench(17920, 5120, 1280)
bench(71680, 1280, 1280)
