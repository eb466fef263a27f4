"""GPU bring-up of the tcgen05 implicit GEMM: prints max errors vs torch fp32 references."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ctrlv_b200 import ops

torch.manual_seed(0)
dev = "cuda"
BF = torch.bfloat16

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()

def report(name, got, ref):
    r, m = rel(got, ref)
    ok = r < 1e-2
    print(f"{'OK  ' if ok else 'FAIL'} {name}: rel_l2={r:.3e} max_abs={m:.3e} |ref|max={ref.abs().max().item():.3e}", flush=True)
    return ok

def test_linear(M, K, N, **flags):
    a = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    ref = a.float() @ w.float().t()
    kw = {}
    if flags.get("bias"):
        b = torch.randn(N, device=dev); kw["bias"] = b; ref = ref + b
    if flags.get("rowbias"):
        R = 4; rb = torch.randn(R, N, device=dev)
        kw.update(rowbias=rb, rb_mode=2, rb_div=max(M // 8, 1), rb_mod=R)
        idx = (torch.arange(M, device=dev) // max(M // 8, 1)) % R
        ref = ref + rb[idx]
    if flags.get("geglu"):
        kw["geglu"] = True
        ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    s_acc = flags.get("s_acc", 1.0)
    ref = ref * s_acc; kw["s_acc"] = s_acc
    No = ref.shape[1]
    if flags.get("res1"):
        r1 = torch.randn(M, No, device=dev).to(BF); kw.update(res1=r1, s_res1=0.7); ref = ref + 0.7 * r1.float()
    if flags.get("res2"):
        r2 = torch.randn(M, No, device=dev).to(BF); kw.update(res2=r2, s_res2=0.3); ref = ref + 0.3 * r2.float()
    out = ops.linear(a, w, **kw)
    torch.cuda.synchronize()
    return report(f"linear M={M} K={K} N={N} {flags}", out, ref)

def pack_conv_w(w):  # [Cout, Cin, 3, 3] -> [Cout, 9*Cin] tap-major
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()

def test_conv(frames, H, W, Cin, Cout, stride=1, shortcut=False):
    x = torch.randn(frames, Cin, H, W, device=dev).to(BF)
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).to(BF)
    b = torch.randn(Cout, device=dev)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
    xl = x.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous()
    wp = pack_conv_w(w)
    kw = dict(bias=b)
    if shortcut:
        ws = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).to(BF)
        x2 = torch.randn(frames, Cin, H, W, device=dev).to(BF)
        ref = ref + F.conv2d(x2.float(), ws.float()[:, :, None, None])
        wp = torch.cat([wp, ws], dim=1).contiguous()
        kw["sc0"] = x2.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous()
    out = ops.conv3x3(xl, frames, H, W, wp, stride=stride, **kw)
    torch.cuda.synchronize()
    refl = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    return report(f"conv3x3 f={frames} {H}x{W} {Cin}->{Cout} s={stride} sc={shortcut}", out, refl)

def test_tconv(B, T, HW, Cc, Cout):
    x = torch.randn(B, Cc, T, HW, 1, device=dev).to(BF)
    w = (torch.randn(Cout, Cc, 3, 1, 1, device=dev) / (3 * Cc) ** 0.5).to(BF)
    b = torch.randn(Cout, device=dev)
    ref = F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))  # [B, Cout, T, HW, 1]
    xl = x[..., 0].permute(0, 2, 3, 1).reshape(-1, Cc).contiguous()
    wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(Cout, 3 * Cc).contiguous()
    out = ops.conv_t3(xl, B, T, HW, wp, bias=b)
    torch.cuda.synchronize()
    refl = ref[..., 0].permute(0, 2, 3, 1).reshape(-1, Cout)
    return report(f"conv_t3 B={B} T={T} HW={HW} {Cc}->{Cout}", out, refl)

def bench_linear(M, K, N, iters=20, **kw):
    a = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    out = torch.empty(M, N // 2 if kw.get("geglu") else N, device=dev, dtype=BF)
    for _ in range(3): ops.linear(a, w, out=out, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.linear(a, w, out=out, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"bench linear M={M} K={K} N={N} {kw}: {ms*1e3:.1f} us  {2*M*K*N/ms/1e9:.1f} TFLOP/s", flush=True)
    ref = torch.empty(M, N, device=dev, dtype=BF)
    for _ in range(3): torch.matmul(a, w.t(), out=ref)
    e0.record()
    for _ in range(iters): torch.matmul(a, w.t(), out=ref)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"      cuBLAS same shape: {ms*1e3:.1f} us  {2*M*K*N/ms/1e9:.1f} TFLOP/s", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    ok = True
    ok &= test_linear(128, 64, 64)
    ok &= test_linear(256, 128, 128)
    ok &= test_linear(1000, 320, 320, bias=True)
    ok &= test_linear(4096, 320, 960)
    ok &= test_linear(4096, 1280, 1280, bias=True, res1=True)
    ok &= test_linear(1120, 1280, 10240, bias=True, geglu=True)
    ok &= test_linear(2048, 320, 2560, bias=True, geglu=True, rowbias=True)
    ok &= test_linear(2048, 640, 640, bias=True, rowbias=True, res1=True, res2=True, s_acc=0.4)
    ok &= test_conv(2, 16, 16, 64, 64)
    ok &= test_conv(3, 40, 64, 64, 128)
    ok &= test_conv(4, 10, 16, 128, 128)
    ok &= test_conv(5, 5, 8, 128, 64)
    ok &= test_conv(2, 16, 16, 64, 128, shortcut=True)
    ok &= test_conv(2, 16, 16, 64, 64, stride=2)
    ok &= test_conv(3, 40, 64, 128, 128, stride=2)
    ok &= test_tconv(2, 14, 160, 128, 128)
    ok &= test_tconv(2, 14, 40, 64, 64)
    ok &= test_tconv(1, 5, 256, 64, 128)
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if "--bench" in sys.argv:
        bench_linear(71680, 320, 320)
        bench_linear(71680, 320, 2560, geglu=True)
        bench_linear(71680, 1280, 320)
        bench_linear(17920, 640, 5120, geglu=True)
        bench_linear(17920, 2560, 640)
        bench_linear(4480, 1280, 10240, geglu=True)
        bench_linear(8192, 8192, 8192)
