"""Microbenchmark of the step's dominant igemm problems (CUDA events, 256 MB L2 flush between calls).
   python scripts/bench_igemm.py [tag]        -> gpurun_out/igemm_<tag>.json
A/B of two builds on one box: CTRLV_B200_LIB=<other .so> python scripts/bench_igemm.py base"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import ops
BF, dev = torch.bfloat16, "cuda"
tag = sys.argv[1] if len(sys.argv) > 1 else "run"
# STREAMK=1: whole tiles only, 2: stream-K wherever the problem allows it (default 0: the launcher's heuristic)
ops.lib().ctrlv_igemm_streamk(int(os.environ.get("STREAMK", "0")))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=8):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters

def lin(M, K, N, geglu=False, res1=False, res2=False, rowbias=False):
    a = torch.randn(M, K, device=dev).to(BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    b = torch.randn(N, device=dev)
    no = N // 2 if geglu else N
    out = torch.empty(M, no, device=dev, dtype=BF)
    kw = dict(bias=b, out=out, geglu=geglu)
    if res1: kw["res1"] = torch.randn(M, no, device=dev).to(BF)
    if res2: kw["res2"] = torch.randn(M, no, device=dev).to(BF); kw["s_res2"] = 0.5
    if rowbias: kw.update(rowbias=torch.randn(2, no, device=dev), rb_mode=1, rb_div=M // 2)
    return (lambda: ops.linear(a, w, **kw)), 2.0 * M * K * N

def conv(F_, H, W, C, N):
    x = torch.randn(F_ * H * W, C, device=dev).to(BF); w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).to(BF)
    b = torch.randn(N, device=dev); out = torch.empty(F_ * H * W, N, device=dev, dtype=BF)
    return (lambda: ops.conv3x3(x, F_, H, W, w, bias=b, out=out)), 2.0 * F_ * H * W * 9 * C * N

def convt(B, T, HW, C, N):
    x = torch.randn(B * T * HW, C, device=dev).to(BF); w = (torch.randn(N, 3 * C, device=dev) / (3 * C) ** 0.5).to(BF)
    b = torch.randn(N, device=dev); out = torch.empty(B * T * HW, N, device=dev, dtype=BF)
    res = torch.randn(B * T * HW, N, device=dev).to(BF)
    return (lambda: ops.conv_t3(x, B, T, HW, w, bias=b, out=out, res1=res, s_acc=0.5)), 2.0 * B * T * HW * 3 * C * N

def ffln(M, C, fused):
    x = torch.randn(M, C, device=dev).to(BF); n = torch.empty_like(x)
    w1 = (torch.randn(8 * C, C, device=dev) / C ** 0.5).to(BF); b1 = torch.randn(8 * C, device=dev)
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(BF); b2 = torch.randn(C, device=dev)
    out = torch.empty(M, C, device=dev, dtype=BF)
    if fused:
        return (lambda: ops.feedforward(x, w1, b1, w2, ln_eps=1e-5, bias=b2, out=out, res1=x)), 2.0 * M * C * 12 * C
    return (lambda: ops.feedforward(ops.layernorm(x, out=n), w1, b1, w2, bias=b2, out=out, res1=x)), 2.0 * M * C * 12 * C

def ff(M, C, fused, res2=False):
    x = torch.randn(M, C, device=dev).to(BF)
    w1 = (torch.randn(8 * C, C, device=dev) / C ** 0.5).to(BF); b1 = torch.randn(8 * C, device=dev)
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(BF); b2 = torch.randn(C, device=dev)
    out = torch.empty(M, C, device=dev, dtype=BF); hid = torch.empty(M, 4 * C, device=dev, dtype=BF)
    kw = dict(bias=b2, out=out, res1=torch.randn(M, C, device=dev).to(BF))
    if res2: kw["res2"] = torch.randn(M, C, device=dev).to(BF); kw["s_res2"] = 0.5
    if fused:
        return (lambda: ops.feedforward(x, w1, b1, w2, **kw)), 2.0 * M * C * 12 * C
    return (lambda: ops.linear(ops.linear(x, w1, bias=b1, geglu=True, out=hid), w2, **kw)), 2.0 * M * C * 12 * C

def lnlin(M, K, N, fused):
    x = torch.randn(M, K, device=dev).to(BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    b = torch.randn(N, device=dev); out = torch.empty(M, N, device=dev, dtype=BF); n = torch.empty_like(x)
    if fused:
        return (lambda: ops.linear_ln(x, w, bias=b, out=out)), 2.0 * M * K * N
    return (lambda: ops.linear(ops.layernorm(x, out=n), w, bias=b, out=out)), 2.0 * M * K * N

CASES = [
    ("ln+qkv L0 one launch 320->960", lambda: lnlin(71680, 320, 960, True), 14),
    ("ln+qkv L0 two launches 320->960", lambda: lnlin(71680, 320, 960, False), 0),
    ("ln+ff L0 one launch 71680x320", lambda: ffln(71680, 320, True), 21),
    ("ln+ff L0 two launches 71680x320", lambda: ffln(71680, 320, False), 0),
    ("ff L0 fused 71680x320", lambda: ff(71680, 320, True), 14),
    ("ff L0 fused 71680x320 +res2", lambda: ff(71680, 320, True, True), 7),
    ("ff L0 two-launch 71680x320", lambda: ff(71680, 320, False), 0),
    ("geglu_up L0 71680x320->2560", lambda: lin(71680, 320, 2560, geglu=True), 21),
    ("geglu_up L1 17920x640->5120", lambda: lin(17920, 640, 5120, geglu=True), 21),
    ("geglu_up L2 4480x1280->10240", lambda: lin(4480, 1280, 10240, geglu=True), 21),
    ("geglu_up L3 1120x1280->10240", lambda: lin(1120, 1280, 10240, geglu=True), 6),
    ("lin L0 320->320 +res", lambda: lin(71680, 320, 320, res1=True, rowbias=True), 21),
    ("lin L0 320->320", lambda: lin(71680, 320, 320), 10),
    ("qkv L0 320->960", lambda: lin(71680, 320, 960), 14),
    ("down L0 1280->320 +res", lambda: lin(71680, 1280, 320, res1=True), 14),
    ("down L0 1280->320 +res+res2", lambda: lin(71680, 1280, 320, res1=True, res2=True), 7),
    ("qkv L1 640->1920", lambda: lin(17920, 640, 1920), 14),
    ("down L1 2560->640 +res", lambda: lin(17920, 2560, 640, res1=True), 14),
    ("lin L1 640->640 +res", lambda: lin(17920, 640, 640, res1=True, rowbias=True), 21),
    ("down L2 5120->1280 +res", lambda: lin(4480, 5120, 1280, res1=True), 14),
    ("qkv L2 1280->3840", lambda: lin(4480, 1280, 3840), 14),
    ("lin L2 1280->1280 +res", lambda: lin(4480, 1280, 1280, res1=True, rowbias=True), 21),
    ("lin L3 1280->1280 +res", lambda: lin(1120, 1280, 1280, res1=True), 6),
    ("down L3 5120->1280 +res", lambda: lin(1120, 5120, 1280, res1=True), 6),
    ("qkv L3 1280->3840", lambda: lin(1120, 1280, 3840), 4),
    ("conv3x3 L0 320->320", lambda: conv(28, 40, 64, 320, 320), 8),
    ("conv3x3 L1 640->640", lambda: conv(28, 20, 32, 640, 640), 4),
    ("conv3x3 L2 1280->1280", lambda: conv(28, 10, 16, 1280, 1280), 5),
    ("conv3x3 L3 1280->1280", lambda: conv(28, 5, 8, 1280, 1280), 16),
    ("conv3x3 L3 2560->1280", lambda: conv(28, 5, 8, 2560, 1280), 3),
    ("conv_t3 L0 320", lambda: convt(2, 14, 2560, 320, 320), 14),
    ("conv_t3 L1 640", lambda: convt(2, 14, 640, 640, 640), 14),
    ("conv_t3 L2 1280", lambda: convt(2, 14, 160, 1280, 1280), 14),
    ("conv_t3 L3 1280", lambda: convt(2, 14, 40, 1280, 1280), 22),
]
only = os.environ.get("CASES")
rows, tot = [], 0.0
for name, mk, n in CASES:
    if only and not any(o in name for o in only.split(",")): continue
    fn, fl = mk()
    ms = timeit(fn)
    rows.append({"case": name, "us": round(ms * 1e3, 1), "tflops": round(fl / ms / 1e9, 1), "per_step": n, "ms_per_step": round(ms * n, 3)})
    tot += ms * n
    print(json.dumps(rows[-1]), flush=True)
    del fn
    torch.cuda.empty_cache()
print(json.dumps({"tag": tag, "sum_ms_per_step": round(tot, 3)}), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"tag": tag, "rows": rows, "sum_ms_per_step": tot}, open(f"gpurun_out/igemm_{tag}.json", "w"), indent=1)
