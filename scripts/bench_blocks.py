"""BASELINE config 5: per-block microbenchmark sweep (temporal attention, spatial attention,
SpatioTemporalResBlock, transformer block) at 320x512 (T=14) and 576x1024 (T=25) vs roofline."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, ops
BF = torch.bfloat16; dev = "cuda"
PEAK_TF = 1398.3; PEAK_GBS = 6551.0
try:
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    PEAK_TF = pk["bf16_tflops_sustained"]; PEAK_GBS = pk["hbm_gbs"]
except Exception:
    pass
def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
rows = []
def rec(name, ms, flops=None, bytes_=None):
    r = {"block": name, "us": round(ms * 1e3, 1)}
    if flops: r["tflops"] = round(flops / ms / 1e9, 1); r["frac_tensor_peak"] = round(flops / ms / 1e9 / PEAK_TF, 3)
    if bytes_: r["gbs"] = round(bytes_ / ms / 1e6, 1); r["frac_hbm_peak"] = round(bytes_ / ms / 1e6 / PEAK_GBS, 3)
    rows.append(r); print(json.dumps(r), flush=True)
sd = models.random_state_dict(dict(models.SVD_CONFIG), False, seed=0)
for (T, h, w) in ((14, 40, 64), (25, 72, 128)):
    B = 2; tag = f"T{T}_{h}x{w}"
    for lvl, (C, heads) in enumerate(((320, 5), (640, 10), (1280, 20))):
        hh, ww = h >> lvl, w >> lvl; S = hh * ww; F_ = B * T
        qkv = torch.randn(F_ * S, 3 * C, device=dev).to(BF); out = torch.empty(F_ * S, C, device=dev, dtype=BF)
        ms = timeit(lambda: ops.attn_spatial(qkv, F_, S, heads, out=out))
        rec(f"{tag} spatial_attn S={S} heads={heads}", ms, flops=4.0 * F_ * heads * S * S * 64, bytes_=F_ * S * C * 8)
        ms = timeit(lambda: ops.attn_temporal(qkv, B, T, S, heads, out=out))
        rec(f"{tag} temporal_attn S={S} heads={heads}", ms, flops=4.0 * B * S * heads * T * T * 64, bytes_=F_ * S * C * 8)
        del qkv, out
    # one SpatioTemporalResBlock and one transformer per level (down-path weights of the real config)
    emb = torch.randn(B, 1280, device=dev); ehs = torch.randn(B, 1024, device=dev)
    for lvl, (pfx_r, pfx_a, C, heads) in enumerate((("down_blocks.0.resnets.1", "down_blocks.0.attentions.1", 320, 5),
                                                    ("down_blocks.1.resnets.1", "down_blocks.1.attentions.1", 640, 10),
                                                    ("down_blocks.2.resnets.1", "down_blocks.2.attentions.1", 1280, 20))):
        hh, ww = h >> lvl, w >> lvl; M = B * T * hh * ww; g = (B, T, hh, ww)
        rb = models._ResBlock(sd, pfx_r, 1e-6); tr = models._Transformer(sd, pfx_a, heads, "s_major")
        holder = type("H", (), {})()
        holder.temb_w = torch.cat([rb.temb.w, rb.ttemb.w]).contiguous(); holder.temb_b = torch.cat([rb.temb.b, rb.ttemb.b]).contiguous()
        rb.temb_off, rb.ttemb_off = 0, C
        cw = [tr.attn2.w, tr.tattn2.w]; cb = [tr.attn2.b, tr.tattn2.b]
        tr.attn2.off, tr.tattn2.off = 0, C
        ctx_w = models._w(torch.cat(cw)); ctx_b = models._f(torch.cat(cb))
        from types import SimpleNamespace
        ctx = ops.small_linear(ehs, ctx_w, ctx_b)
        aux = SimpleNamespace(temb=ops.small_linear(emb, holder.temb_w, holder.temb_b, act_in=True), ctx=ctx,
                              ctx_all=ctx, vB=ctx.shape[0], b0=0)
        x = torch.randn(M, C, device=dev).to(BF)
        ms = timeit(lambda: rb(x, aux, g))
        fl = 2.0 * M * (18 * C * C + 6 * C * C)
        rec(f"{tag} SpatioTemporalResBlock C={C} {hh}x{ww}", ms, flops=fl)
        ms = timeit(lambda: tr(x, aux, g))
        S = hh * ww
        fl = 2.0 * M * C * C * 14 + 3 * 2.0 * M * 12 * C * C + 4.0 * B * T * heads * S * S * 64 + 4.0 * B * S * heads * T * T * 64
        rec(f"{tag} TransformerSpatioTemporal C={C} S={S}", ms, flops=fl)
        del x, rb, tr
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_blocks.json", "w"), indent=1)
