import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTRLV_DEBUG_TRACE"] = "1"
import torch
from ctrlv_b200 import ops, _lib
BF = torch.bfloat16
def run(M, K, N, res=False, geglu=False, tag=""):
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF); b = torch.randn(N, device="cuda")
    kw = dict(bias=b)
    if res: kw["res1"] = torch.randn(M, N, device="cuda").to(BF)
    if geglu: kw["geglu"] = True
    out = torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=BF)
    for _ in range(3): ops.linear(a, w, out=out, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.linear(a, w, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib = _lib.load(); lib.ctrlv_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.ctrlv_debug_trace_read(buf, 512)
    t = [buf[i] for i in range(512)]
    t0 = min(x for x in t[:8] if x > 0)
    print(f"--- {tag} M={M} K={K} N={N} res={res} geglu={geglu}: {e0.elapsed_time(e1)*100:.1f} us")
    for it in (2, 3, 4):
        row = [t[it * 8 + k] - t0 if t[it * 8 + k] else -1 for k in range(7)]
        print("   tile", it, row, " mainloop", row[3] - row[2], " epilogue", row[6] - row[5])
tag = f"BRES={os.environ.get('CTRLV_DEBUG_BRES')} BN={os.environ.get('CTRLV_DEBUG_BN')}"
run(71680, 320, 320, tag=tag)
run(71680, 320, 320, res=True, tag=tag)
run(71680, 320, 960, tag=tag)
run(71680, 320, 2560, geglu=True, tag=tag)
run(71680, 1280, 320, res=True, tag=tag)
