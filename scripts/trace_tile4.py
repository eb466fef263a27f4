import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTRLV_DEBUG_TRACE"] = "1"
import torch
from ctrlv_b200 import ops, _lib
BF = torch.bfloat16
M, K, N = 71680, 1280, 1280
a = torch.randn(M, K, device="cuda").to(BF); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
out = torch.empty(M, N, device="cuda", dtype=BF)
for _ in range(3): ops.linear(a, w, out=out)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib = _lib.load(); lib.ctrlv_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.ctrlv_debug_trace_read(buf, 512)
t = [buf[i] for i in range(512)]
base = t[40 * 8 + 0]
print(f"BN={os.environ.get('CTRLV_DEBUG_BN')}: tile 3, k-blocks 8..11: [mma: before wait, after wait, after 4 mma, after commit | producer: before empty-wait, after, after tma issue]")
for j in range(4):
    print([t[(40 + j) * 8 + k] - base for k in range(7)])
