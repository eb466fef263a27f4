import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTRLV_DEBUG_TRACE"] = "1"
import torch
from ctrlv_b200 import ops, _lib
BF = torch.bfloat16
def run(M, K, N, tag=""):
    a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
    out = torch.empty(M, N, device="cuda", dtype=BF)
    for _ in range(3): ops.linear(a, w, out=out)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.linear(a, w, out=out)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib = _lib.load(); lib.ctrlv_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    lib.ctrlv_debug_trace_read(buf, 512)
    t = [buf[i] for i in range(512)]
    ml = [t[it * 8 + 3] - t[it * 8 + 2] for it in (2, 3, 4, 5)]
    ep = [t[it * 8 + 6] - t[it * 8 + 5] for it in (2, 3, 4, 5)]
    per = [t[(it + 1) * 8 + 2] - t[it * 8 + 2] for it in (2, 3, 4)]
    print(f"{tag} M={M} K={K} N={N}: {e0.elapsed_time(e1)*100:.1f} us {2*M*K*N/e0.elapsed_time(e1)/1e7:.0f} TF | mainloop {ml} per-kblock {[m // (K // 64 - 1) for m in ml]} | epilogue {ep} | tile period {per}", flush=True)
tag = f"BN={os.environ.get('CTRLV_DEBUG_BN')} ST={os.environ.get('CTRLV_DEBUG_STAGES')}"
run(71680, 320, 1280, tag)
run(17920, 320, 1280, tag)
run(71680, 1280, 1280, tag)
run(17920, 5120, 1280, tag)
