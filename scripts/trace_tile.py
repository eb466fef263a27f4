import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTRLV_DEBUG_TRACE"] = "1"
import torch
from ctrlv_b200 import ops, _lib
BF = torch.bfloat16
def run(M, K, N, res=False, geglu=False):
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF); b = torch.randn(N, device="cuda")
    kw = dict(bias=b)
    if res: kw["res1"] = torch.randn(M, N, device="cuda").to(BF)
    if geglu: kw["geglu"] = True
    for _ in range(3): ops.linear(a, w, **kw)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib = _lib.load(); lib.ctrlv_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.ctrlv_debug_trace_read(buf, 512)
    t = [buf[i] for i in range(512)]
    t0 = min(x for x in t[:8] if x > 0)
    print(f"--- M={M} K={K} N={N} res={res} geglu={geglu}: per tile [prod_first_slot, mma_tempty_ok, mma_first_full, mma_last_full, epi_bias_done, epi_tfull_ok, epi_done] (cycles from start)")
    for it in range(10):
        row = [t[it * 8 + k] - t0 if t[it * 8 + k] else -1 for k in range(7)]
        print(it, row)
run(71680, 320, 320)
run(71680, 320, 320, res=True)
run(71680, 320, 2560, geglu=True)
run(71680, 1280, 320, res=True)
