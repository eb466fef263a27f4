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
    base = t[40 * 8 + 6]
    print(f"--- M={M} K={K} N={N} res={res} geglu={geglu}: warp2 per tile: [tile_top, after_bias_barrier(=enter ep_tile), after 1st res issue.., tfull_ok, chunk0 done, chunk1 done, exit]")
    for it in range(2, 7):
        g = lambda k: t[(40 + it) * 8 + k] - base
        print(f"  tile {it}: top {g(6)}  enter {g(4)}  tfull_ok {g(1)}  c0_done {g(2)}  c1_done {g(3)}  exit {g(5)}   | mma: tempty_ok {t[it*8+1]-base} first_full {t[it*8+2]-base} last_full {t[it*8+3]-base}")
run(71680, 320, 320)
run(71680, 320, 320, res=True)
run(71680, 320, 2560, geglu=True)
