"""GPU bring-up of attention / norm / glue kernels vs torch fp32 references."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from ctrlv_b200 import ops

torch.manual_seed(0)
dev = "cuda"; BF = torch.bfloat16

def report(name, got, ref, tol=1e-2):
    got = got.float(); ref = ref.float()
    r = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    nan = torch.isnan(got).any().item()
    ok = (r < tol) and not nan
    print(f"{'OK  ' if ok else 'FAIL'} {name}: rel_l2={r:.3e} max_abs={(got-ref).abs().max().item():.3e} nan={nan}", flush=True)
    return ok

def t_attn_spatial(frames, S, heads, sc=1.0):
    C = heads * 64
    qkv = (torch.randn(frames * S, 3 * C, device=dev) * sc).to(BF)
    out = ops.attn_spatial(qkv, frames, S, heads)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(frames, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(frames * S, C)
    return report(f"attn_spatial f={frames} S={S} heads={heads} sc={sc}", out, ref)

def t_attn_temporal(B, T, S, heads, sc=1.0):
    C = heads * 64
    qkv = (torch.randn(B * T * S, 3 * C, device=dev) * sc).to(BF)
    out = ops.attn_temporal(qkv, B, T, S, heads)
    torch.cuda.synchronize()
    x = qkv.float().view(B, T, S, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)  # [3, B, S, heads, T, 64]
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2])  # [B, S, heads, T, 64]
    ref = ref.permute(0, 3, 1, 2, 4).reshape(B * T * S, C)
    return report(f"attn_temporal B={B} T={T} S={S} heads={heads} sc={sc}", out, ref)

def t_gn(n_units, rpu, C0, C1, silu, eps=1e-5):
    C = C0 + C1
    x0 = (torch.randn(n_units * rpu, C0, device=dev) * 2 + 0.5).to(BF)
    x1 = (torch.randn(n_units * rpu, C1, device=dev) - 0.3).to(BF) if C1 else None
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    out = ops.groupnorm(x0, n_units, rpu, g, b, eps, silu, src1=x1)
    torch.cuda.synchronize()
    x = torch.cat([x0, x1], 1) if C1 else x0
    xr = x.float().view(n_units, rpu, C).permute(0, 2, 1)  # [N, C, L]
    ref = F.group_norm(xr, 32, g, b, eps)
    if silu: ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(-1, C)
    return report(f"groupnorm units={n_units} rows={rpu} C={C0}+{C1} silu={silu}", out, ref)

def t_ln(M, C, rb=False):
    x = (torch.randn(M, C, device=dev) * 1.5 + 0.2).to(BF)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    kw = {}
    xr = x.float()
    if rb:
        T = 7; S = M // (2 * T)
        pos = torch.randn(T, C, device=dev)
        kw = dict(rowbias=pos, rb_div=S, rb_mod=T)
        idx = (torch.arange(M, device=dev) // S) % T
        xr = xr + pos[idx]
    out = ops.layernorm(x, g, b, 1e-5, **kw)
    torch.cuda.synchronize()
    ref = F.layer_norm(xr, (C,), g, b, 1e-5)
    return report(f"layernorm M={M} C={C} rb={rb}", out, ref)

def t_small_linear(M, K, N, act_in, act_out):
    x = torch.randn(M, K, device=dev); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF); b = torch.randn(N, device=dev)
    out = ops.small_linear(x, w, b, act_in, act_out)
    torch.cuda.synchronize()
    xi = F.silu(x) if act_in else x
    ref = xi @ w.float().t() + b
    if act_out: ref = F.silu(ref)
    return report(f"small_linear M={M} K={K} N={N} {act_in} {act_out}", out, ref, 1e-4)

def t_sinusoid():
    t = torch.tensor([1.63777006, -1.55365205, 6.0, 127.0, 0.02, 13.0], device=dev)
    out = ops.sinusoid(t, 320, round_bf16=False)
    import math
    half = 160
    f = torch.exp(-math.log(10000) * torch.arange(half, device=dev, dtype=torch.float32) / half)
    a = t[:, None] * f[None]
    ref = torch.cat([torch.cos(a), torch.sin(a)], -1)
    print("   sinusoid KAT t=1.63777006:", out[0, :3].tolist(), out[0, 160:163].tolist())
    return report("sinusoid", out, ref, 1e-5)

def t_glue():
    ok = True
    B, T, h, w = 2, 3, 8, 16
    lat = torch.randn(B, T, 4, h, w, device=dev) * 10
    img = torch.randn(2 * B, T, 4, h, w, device=dev); ctl = torch.randn(2 * B, T, 4, h, w, device=dev)
    sigma = 3.7
    sd = torch.tensor([sigma, 2.9], device=dev)
    out = ops.prep_input(lat, img, ctl, True, sd)
    ref = torch.zeros(2 * B, T, h, w, 64, device=dev)
    ref[..., 0:4] = (torch.cat([lat, lat]) / (sigma ** 2 + 1) ** 0.5).permute(0, 1, 3, 4, 2)
    ref[..., 4:8] = img.permute(0, 1, 3, 4, 2); ref[..., 8:12] = ctl.permute(0, 1, 3, 4, 2)
    ok &= report("prep_input", out, ref.reshape(-1, 64), 5e-3)
    noise = torch.randn(2 * B * T * h * w, 4, device=dev)
    g = torch.linspace(1, 3, T, device=dev)
    lat2 = lat.clone()
    ops.cfg_euler(lat2, noise, True, g, sd)
    n5 = noise.view(2 * B, T, h, w, 4).permute(0, 1, 4, 2, 3)
    nu, nc = n5[:B], n5[B:]
    v = nu + g.view(1, T, 1, 1, 1) * (nc - nu)
    x0 = v * (-sigma / (sigma ** 2 + 1) ** 0.5) + lat / (sigma ** 2 + 1)
    ref = lat + (lat - x0) / sigma * (2.9 - sigma)
    ok &= report("cfg_euler", lat2, ref, 1e-5)
    x = torch.randn(3 * 4 * 6, 64, device=dev).to(BF)
    up = ops.upsample2x(x, 3, 4, 6)
    ref = F.interpolate(x.float().view(3, 4, 6, 64).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).reshape(-1, 64)
    ok &= report("upsample2x", up, ref, 1e-6)
    y = torch.randn_like(x.float()).to(BF)
    ok &= report("axpby", ops.axpby(x, y, 0.3, 0.7), 0.3 * x.float() + 0.7 * y.float(), 5e-3)
    src = torch.randn(5, 8, 4, 6, device=dev)
    buf = torch.zeros(5 * 24, 64, device=dev, dtype=BF)
    ops.nchw_to_nhwc(src, buf, 4)
    ref = torch.zeros(5, 4, 6, 64, device=dev); ref[..., 4:12] = src.permute(0, 2, 3, 1)
    ok &= report("nchw_to_nhwc", buf, ref.reshape(-1, 64), 5e-3)
    back = ops.nhwc_to_nchw(buf[:, 4:12], 5, 8, 4, 6)
    ok &= report("nhwc_to_nchw", back, src, 5e-3)
    return ok

if __name__ == "__main__":
    ok = True
    ok &= t_attn_spatial(2, 128, 1)
    ok &= t_attn_spatial(2, 256, 2)
    ok &= t_attn_spatial(3, 160, 2)
    ok &= t_attn_spatial(2, 40, 4)
    ok &= t_attn_spatial(2, 2560, 5)
    ok &= t_attn_spatial(1, 640, 2, sc=3.0)
    ok &= t_attn_temporal(1, 14, 9, 1)
    ok &= t_attn_temporal(2, 14, 160, 2)
    ok &= t_attn_temporal(2, 25, 64, 2)
    ok &= t_attn_temporal(2, 4, 256, 1, sc=3.0)
    ok &= t_attn_temporal(2, 14, 2560, 5)
    ok &= t_gn(4, 256, 64, 0, True)
    ok &= t_gn(28, 2560, 320, 0, True)
    ok &= t_gn(2, 14 * 640, 640, 0, True)
    ok &= t_gn(6, 160, 1280, 640, True)
    ok &= t_gn(6, 40, 1280, 1280, False, 1e-6)
    ok &= t_gn(3, 640, 640, 320, True)
    ok &= t_ln(4096, 320); ok &= t_ln(1120, 1280); ok &= t_ln(14 * 2 * 24, 64, rb=True); ok &= t_ln(14 * 40, 640, rb=True)
    ok &= t_small_linear(2, 320, 1280, False, True); ok &= t_small_linear(2, 1280, 1280, True, False)
    ok &= t_small_linear(14, 320, 1280, False, True); ok &= t_small_linear(25, 1024, 320, False, False)
    ok &= t_sinusoid()
    ok &= t_glue()
    print("ALL OK" if ok else "SOME FAILED", flush=True)
    if "--bench" in sys.argv:
        def bench(fn, name, flops=None, bytes_=None, iters=20):
            for _ in range(3): fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            s = f"bench {name}: {ms*1e3:.1f} us"
            if flops: s += f"  {flops/ms/1e9:.1f} TFLOP/s"
            if bytes_: s += f"  {bytes_/ms/1e6:.1f} GB/s"
            print(s, flush=True)
        for (f, S, hd) in [(28, 2560, 5), (28, 640, 10), (28, 160, 20), (28, 40, 20)]:
            C = hd * 64
            qkv = torch.randn(f * S, 3 * C, device=dev).to(BF); out = torch.empty(f * S, C, device=dev, dtype=BF)
            bench(lambda: ops.attn_spatial(qkv, f, S, hd, out=out), f"attn_spatial f={f} S={S} heads={hd}", flops=4 * f * hd * S * S * 64, bytes_=f * S * C * 8)
            q, k, v = qkv.view(f, S, 3, hd, 64).permute(2, 0, 3, 1, 4)
            bench(lambda: F.scaled_dot_product_attention(q, k, v), "   torch SDPA same", flops=4 * f * hd * S * S * 64)
        for (S, hd) in [(2560, 5), (640, 10), (160, 20)]:
            C = hd * 64
            qkv = torch.randn(2 * 14 * S, 3 * C, device=dev).to(BF); out = torch.empty(2 * 14 * S, C, device=dev, dtype=BF)
            bench(lambda: ops.attn_temporal(qkv, 2, 14, S, hd, out=out), f"attn_temporal S={S} heads={hd}", bytes_=2 * 14 * S * C * 8)
        x = torch.randn(71680, 320, device=dev).to(BF); g = torch.randn(320, device=dev); b = torch.randn(320, device=dev); o = torch.empty_like(x)
        bench(lambda: ops.groupnorm(x, 28, 2560, g, b, 1e-5, True, out=o), "groupnorm L0 spatial", bytes_=71680 * 320 * 6)
        bench(lambda: ops.groupnorm(x, 2, 35840, g, b, 1e-5, True, out=o), "groupnorm L0 temporal", bytes_=71680 * 320 * 6)
        bench(lambda: ops.layernorm(x, g, b, out=o), "layernorm L0", bytes_=71680 * 320 * 4)
