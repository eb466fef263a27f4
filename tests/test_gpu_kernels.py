"""GPU parity of every sm_100a kernel, called through the C ABI (ctypes), against the oracle's
ops (torch fp32 on the same bf16-rounded inputs).  Tolerances: the kernels accumulate in fp32 and
round once to bf16 on store (relative 2^-9 per element => rel-L2 ~ 1.7e-3); fp32 outputs 1e-5."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
TOL_BF16 = 4e-3   # rel-L2 for bf16-stored outputs
TOL_F32 = 2e-5


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def ops():
    from ctrlv_b200 import _lib, ops as o
    assert _lib.load().ctrlv_device_check() == 0, _lib.load().ctrlv_last_error()
    torch.manual_seed(0)
    return o


@pytest.mark.parametrize("M,K,N,flags", [
    (128, 64, 64, {}),
    (1, 64, 32, {}),                                # single row, narrowest tile
    (1000, 320, 320, dict(bias=True)),              # ragged M
    (4096, 320, 960, {}),
    (4096, 1280, 1280, dict(bias=True, res1=True)),
    (1120, 1280, 10240, dict(bias=True, geglu=True)),
    (2048, 320, 2560, dict(bias=True, geglu=True, rowbias=2)),
    (2048, 640, 640, dict(bias=True, rowbias=2, res1=True, res2=True, s_acc=0.4)),
    (2 * 3 * 50, 128, 128, dict(bias=True, rowbias=3, res1=True)),   # time_context quirk indexing
    (71680, 320, 320, dict(bias=True, res1=True)),  # BASELINE config-2 level-0 shape
])
def test_linear(ops, M, K, N, flags):
    dev = "cuda"
    a = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    ref = a.float() @ w.float().t()
    kw = {}
    if flags.get("bias"):
        b = torch.randn(N, device=dev); kw["bias"] = b; ref = ref + b
    mode = flags.get("rowbias", 0)
    if mode == 2:
        R = 4; rb = torch.randn(R, N, device=dev); div = max(M // 8, 1)
        kw.update(rowbias=rb, rb_mode=2, rb_div=div, rb_mod=R)
        ref = ref + rb[(torch.arange(M, device=dev) // div) % R]
    elif mode == 3:
        B, T, S = 2, 3, M // 6
        rb = torch.randn(B, N, device=dev)
        kw.update(rowbias=rb, rb_mode=3, rb_div=T * S, rb_mod=S, rb_B=B)
        m = torch.arange(M, device=dev)
        ref = ref + rb[((m // (T * S)) * S + m % S) % B]
    if flags.get("geglu"):
        kw["geglu"] = True
        ref = ref[:, 0::2] * F.gelu(ref[:, 1::2])
    s_acc = flags.get("s_acc", 1.0)
    ref = ref * s_acc; kw["s_acc"] = s_acc
    if flags.get("res1"):
        r1 = torch.randn(M, ref.shape[1], device=dev).to(BF); kw.update(res1=r1, s_res1=0.7); ref = ref + 0.7 * r1.float()
    if flags.get("res2"):
        r2 = torch.randn(M, ref.shape[1], device=dev).to(BF); kw.update(res2=r2, s_res2=0.3); ref = ref + 0.3 * r2.float()
    out = ops.linear(a, w, **kw)
    torch.cuda.synchronize()
    assert out.shape == ref.shape and not torch.isnan(out).any()
    assert rel(out, ref) < TOL_BF16


def test_linear_fp32_output_and_padded_columns(ops):
    a = torch.randn(300, 128, device="cuda").to(BF)
    w = torch.zeros(32, 128, device="cuda"); w[:4] = torch.randn(4, 128, device="cuda") / 11
    b = torch.zeros(32, device="cuda"); b[:4] = torch.randn(4, device="cuda")
    out = torch.full((300, 4), 7.0, device="cuda")
    ops.linear(a, w.to(BF), bias=b, out_f32=out, n_store=4)
    torch.cuda.synchronize()
    ref = a.float() @ w.to(BF).float().t()[:, :4] + b[:4]
    assert rel(out, ref) < TOL_F32


def _pack9(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


@pytest.mark.parametrize("frames,H,W,Cin,Cout,stride,shortcut", [
    (2, 16, 16, 64, 64, 1, False),
    (3, 40, 64, 64, 128, 1, False),
    (4, 10, 16, 128, 128, 1, False),    # box packs several frames per tile
    (5, 5, 8, 128, 64, 1, False),       # 120-row tiles
    (2, 16, 16, 64, 128, 1, True),      # fused 1x1 conv_shortcut in the K loop
    (2, 16, 16, 64, 64, 2, False),      # Downsample2D
    (3, 40, 64, 128, 128, 2, False),
    (1, 8, 8, 64, 64, 1, False),
])
def test_conv3x3(ops, frames, H, W, Cin, Cout, stride, shortcut):
    dev = "cuda"
    x = torch.randn(frames, Cin, H, W, device=dev).to(BF)
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).to(BF)
    b = torch.randn(Cout, device=dev)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
    xl = x.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous()
    wp = _pack9(w)
    kw = dict(bias=b)
    if shortcut:
        ws = (torch.randn(Cout, Cin, device=dev) / Cin ** 0.5).to(BF)
        x2 = torch.randn(frames, Cin, H, W, device=dev).to(BF)
        ref = ref + F.conv2d(x2.float(), ws.float()[:, :, None, None])
        wp = torch.cat([wp, ws], dim=1).contiguous()
        kw["sc0"] = x2.permute(0, 2, 3, 1).reshape(-1, Cin).contiguous()
    out = ops.conv3x3(xl, frames, H, W, wp, stride=stride, **kw)
    torch.cuda.synchronize()
    assert rel(out, ref.permute(0, 2, 3, 1).reshape(-1, Cout)) < TOL_BF16


def test_conv3x3_two_sources_is_channel_concat(ops):
    dev = "cuda"
    f, H, W, C0, C1, Co = 2, 8, 16, 64, 128, 64
    x0 = torch.randn(f, C0, H, W, device=dev).to(BF); x1 = torch.randn(f, C1, H, W, device=dev).to(BF)
    w = (torch.randn(Co, C0 + C1, 3, 3, device=dev) / 40).to(BF)
    ref = F.conv2d(torch.cat([x0, x1], 1).float(), w.float(), padding=1)
    rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
    out = ops.conv3x3(rows(x0), f, H, W, _pack9(w), src1=rows(x1))
    torch.cuda.synchronize()
    assert rel(out, rows(ref)) < TOL_BF16


@pytest.mark.parametrize("B,T,HW,C,Co", [(2, 14, 160, 128, 128), (2, 14, 40, 64, 64), (1, 5, 256, 64, 128), (2, 1, 64, 64, 64)])
def test_conv_t3(ops, B, T, HW, C, Co):
    dev = "cuda"
    x = torch.randn(B, C, T, HW, 1, device=dev).to(BF)
    w = (torch.randn(Co, C, 3, 1, 1, device=dev) / (3 * C) ** 0.5).to(BF)
    b = torch.randn(Co, device=dev)
    ref = F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))[..., 0].permute(0, 2, 3, 1).reshape(-1, Co)
    xl = x[..., 0].permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(Co, 3 * C).contiguous()
    out = ops.conv_t3(xl, B, T, HW, wp, bias=b)
    torch.cuda.synchronize()
    assert rel(out, ref) < TOL_BF16


@pytest.mark.parametrize("frames,S,heads,sc", [(2, 128, 1, 1.0), (3, 160, 2, 1.0), (2, 40, 4, 1.0), (2, 2560, 5, 1.0),
                                               (1, 640, 2, 3.0), (1, 1, 1, 1.0), (2, 129, 1, 2.0)])
def test_attn_spatial(ops, frames, S, heads, sc):
    C = heads * 64
    qkv = (torch.randn(frames * S, 3 * C, device="cuda") * sc).to(BF)
    out = ops.attn_spatial(qkv, frames, S, heads)
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(frames, S, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(frames * S, C)
    assert not torch.isnan(out).any() and rel(out, ref) < 5e-3


@pytest.mark.parametrize("B,T,S,heads,sc", [(1, 14, 9, 1, 1.0), (2, 14, 160, 2, 1.0), (2, 25, 64, 2, 1.0),
                                            (2, 4, 256, 1, 3.0), (1, 1, 40, 1, 1.0), (2, 14, 2560, 5, 1.0)])
def test_attn_temporal(ops, B, T, S, heads, sc):
    C = heads * 64
    qkv = (torch.randn(B * T * S, 3 * C, device="cuda") * sc).to(BF)
    out = ops.attn_temporal(qkv, B, T, S, heads)
    torch.cuda.synchronize()
    x = qkv.float().view(B, T, S, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)
    ref = F.scaled_dot_product_attention(x[0], x[1], x[2]).permute(0, 3, 1, 2, 4).reshape(B * T * S, C)
    assert not torch.isnan(out).any() and rel(out, ref) < 5e-3


@pytest.mark.parametrize("units,rows,C0,C1,silu,eps", [(4, 256, 64, 0, True, 1e-5), (28, 2560, 320, 0, True, 1e-6),
                                                       (2, 8960, 640, 0, True, 1e-5), (6, 160, 1280, 640, True, 1e-6),
                                                       (6, 40, 1280, 1280, False, 1e-6), (3, 640, 640, 320, True, 1e-5),
                                                       (2, 4, 64, 0, True, 1e-5),
                                                       # many small units, two sources, ragged row counts
                                                       (8, 40, 1280, 1280, False, 1e-6), (12, 160, 640, 320, True, 1e-5),
                                                       (28, 640, 640, 0, True, 1e-6), (9, 7, 64, 0, True, 1e-5),
                                                       (16, 2561, 128, 0, True, 1e-6)])
def test_groupnorm(ops, units, rows, C0, C1, silu, eps):
    dev = "cuda"
    C = C0 + C1
    x0 = (torch.randn(units * rows, C0, device=dev) * 2 + 0.5).to(BF)
    x1 = (torch.randn(units * rows, C1, device=dev) - 0.3).to(BF) if C1 else None
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    out = ops.groupnorm(x0, units, rows, g, b, eps, silu, src1=x1)
    out2 = ops.groupnorm(x0, units, rows, g, b, eps, silu, src1=x1)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)  # deterministic reduction
    x = torch.cat([x0, x1], 1) if C1 else x0
    ref = F.group_norm(x.float().view(units, rows, C).permute(0, 2, 1), 32, g, b, eps)
    if silu:
        ref = F.silu(ref)
    assert rel(out, ref.permute(0, 2, 1).reshape(-1, C)) < TOL_BF16


@pytest.mark.parametrize("M,C,flags", [
    (128, 64, {}),
    (1000, 64, dict(res1=True)),                                  # ragged M, one hidden chunk ring wrap
    (4096, 128, dict(res1=True, rowbias=True)),                    # tff_in: + frame-position embedding
    (2048, 256, dict(res1=True, res2=True, s_acc=0.3)),            # tff: AlphaBlender epilogue
    (300, 320, dict(res1=True)),                                   # C = 320: two 160-wide MMA2 tiles, 512 TMEM columns
    (71680, 320, dict(res1=True)),                                 # BASELINE config-2 level-0 shape (4 waves of tiles)
])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_feedforward_fused(ops, M, C, flags, cta_group):
    """ctrlv_feedforward (GEGLU up-projection + down-projection in one launch, intermediate in TMEM) against
    the two-launch form and against torch (diffusers FeedForward with exact-erf GELU)."""
    dev = "cuda"
    x = torch.randn(M, C, device=dev).to(BF)
    w1 = (torch.randn(8 * C, C, device=dev) / C ** 0.5).to(BF)      # rows: (value, gate) interleaved
    b1 = torch.randn(8 * C, device=dev) * 0.5
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(BF)
    b2 = torch.randn(C, device=dev)
    kw = dict(bias=b2)
    if flags.get("res1"):
        kw["res1"] = torch.randn(M, C, device=dev).to(BF)
        kw["s_res1"] = 1.0 - flags.get("s_acc", 0.0) if "s_acc" in flags else 1.0
    if flags.get("res2"):
        kw["res2"] = torch.randn(M, C, device=dev).to(BF)
        kw["s_res2"] = 0.7
    if "s_acc" in flags:
        kw["s_acc"] = flags["s_acc"]
    if flags.get("rowbias"):
        T, S = 4, M // 16
        kw.update(rowbias=torch.randn(T, C, device=dev), rb_mode=2, rb_div=S, rb_mod=T)
    from ctrlv_b200 import _lib
    _lib.check(_lib.load().ctrlv_feedforward_override(cta_group))  # single CTAs / CTA pairs (cta_group::2)
    try:
        fused = ops.feedforward(x, w1, b1, w2, **kw)
        fused2 = ops.feedforward(x, w1, b1, w2, **kw)
    finally:
        _lib.load().ctrlv_feedforward_override(0)
    two = ops.linear(ops.linear(x, w1, bias=b1, geglu=True), w2, **kw)
    torch.cuda.synchronize()
    assert torch.equal(fused, fused2)
    # reference: fp32 math on the bf16 operands, hidden activations rounded to bf16 like both kernels do
    u = x.float() @ w1.float().t() + b1
    hid = (u[:, 0::2] * F.gelu(u[:, 1::2])).to(BF).float()
    ref = hid @ w2.float().t() + b2
    if flags.get("rowbias"):
        ref = ref + kw["rowbias"][(torch.arange(M, device=dev) // kw["rb_div"]) % kw["rb_mod"]]
    ref = ref * kw.get("s_acc", 1.0)
    if "res1" in kw:
        ref = ref + kw["s_res1"] * kw["res1"].float()
    if "res2" in kw:
        ref = ref + kw["s_res2"] * kw["res2"].float()
    assert not torch.isnan(fused).any()
    assert rel(fused, ref) < TOL_BF16 and rel(two, ref) < TOL_BF16
    assert rel(fused, two) < 2e-3


@pytest.mark.parametrize("M,C,posemb", [
    (128, 64, False),
    (1000, 128, True),       # ragged M (zero rows past the end normalise to zero), frame-position embedding first
    (300, 320, False),       # C = 320: ten 16-byte pieces per thread, five k-blocks
    (2 * 14 * 160, 320, True),
    (71680, 320, False),     # BASELINE config-2 level-0 shape: four tiles per CTA pair (x re-normalised per tile)
])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_feedforward_fused_with_layernorm(ops, M, C, posemb, cta_group):
    """ctrlv_feedforward_ln (the LayerNorm in front of the FeedForward normalises each x tile in shared memory)
    against ctrlv_layernorm + ctrlv_feedforward and against torch."""
    dev = "cuda"
    x = (torch.randn(M, C, device=dev) * 1.7 + 0.4).to(BF)
    w1 = (torch.randn(8 * C, C, device=dev) / C ** 0.5).to(BF)
    b1 = torch.randn(8 * C, device=dev) * 0.5
    w2 = (torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5).to(BF)
    b2 = torch.randn(C, device=dev)
    kw = dict(bias=b2, res1=x)
    ln = {}
    xin = x.float()
    if posemb:
        T, S = 14, max(M // 28, 1)
        pos = torch.randn(T, C, device=dev)
        ln = dict(rowbias=pos, rb_div=S, rb_mod=T)
        xin = xin + pos[(torch.arange(M, device=dev) // S) % T]
    from ctrlv_b200 import _lib
    _lib.check(_lib.load().ctrlv_feedforward_override(cta_group))
    try:
        fused = ops.feedforward(x, w1, b1, w2, ln_eps=1e-5, ln_rowbias=ln.get("rowbias"), ln_rb_div=ln.get("rb_div", 1),
                                ln_rb_mod=ln.get("rb_mod", 1), **kw)
        fused2 = ops.feedforward(x, w1, b1, w2, ln_eps=1e-5, ln_rowbias=ln.get("rowbias"), ln_rb_div=ln.get("rb_div", 1),
                                 ln_rb_mod=ln.get("rb_mod", 1), **kw)
        n = ops.layernorm(x, **ln)
        two = ops.feedforward(n, w1, b1, w2, **kw)
    finally:
        _lib.load().ctrlv_feedforward_override(0)
    torch.cuda.synchronize()
    assert torch.equal(fused, fused2) and not torch.isnan(fused).any()
    assert rel(fused, two) < 1e-3  # same arithmetic up to the order of the fp32 row sums
    nr = F.layer_norm(xin, (C,), eps=1e-5).to(BF).float()
    u = nr @ w1.float().t() + b1
    hid = (u[:, 0::2] * F.gelu(u[:, 1::2])).to(BF).float()
    ref = hid @ w2.float().t() + b2 + x.float()
    assert rel(fused, ref) < TOL_BF16


@pytest.mark.parametrize("M,K,N,posemb", [
    (128, 64, 64, False),       # one chunk, half of it stored
    (1000, 128, 384, True),     # ragged M, three chunks, two x buffers
    (300, 320, 960, False),     # level-0 q | k | v: 7.5 chunks (the last half chunk is not stored)
    (2 * 14 * 160, 256, 768, False),
    (71680, 320, 960, False),   # BASELINE config-2 level-0 shape
])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_linear_with_layernorm(ops, M, K, N, posemb, cta_group):
    """ctrlv_linear_ln (LayerNorm + Linear in one launch, rows normalised in shared memory) against
    ctrlv_layernorm + ctrlv_linear and against torch."""
    dev = "cuda"
    x = (torch.randn(M, K, device=dev) * 1.3 - 0.2).to(BF)
    w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF)
    b = torch.randn(N, device=dev)
    ln = {}
    xin = x.float()
    if posemb:
        T, S = 5, max(M // 10, 1)
        pos = torch.randn(T, K, device=dev)
        ln = dict(rowbias=pos, rb_div=S, rb_mod=T)
        xin = xin + pos[(torch.arange(M, device=dev) // S) % T]
    from ctrlv_b200 import _lib
    _lib.check(_lib.load().ctrlv_feedforward_override(cta_group))
    try:
        out = torch.full((M, N), 7.0, device=dev, dtype=BF)
        ops.linear_ln(x, w, bias=b, ln_rowbias=ln.get("rowbias"), ln_rb_div=ln.get("rb_div", 1), ln_rb_mod=ln.get("rb_mod", 1), out=out)
        out2 = ops.linear_ln(x, w, bias=b, ln_rowbias=ln.get("rowbias"), ln_rb_div=ln.get("rb_div", 1), ln_rb_mod=ln.get("rb_mod", 1))
        nob = ops.linear_ln(x, w)
    finally:
        _lib.load().ctrlv_feedforward_override(0)
    two = ops.linear(ops.layernorm(x, **ln), w, bias=b)
    torch.cuda.synchronize()
    assert torch.equal(out, out2) and not torch.isnan(out).any()
    assert rel(out, two) < 1e-3
    ref = F.layer_norm(xin, (K,), eps=1e-5).to(BF).float() @ w.float().t()
    assert rel(out, ref + b) < TOL_BF16
    if not posemb:
        assert rel(nob, ref) < TOL_BF16


@pytest.mark.parametrize("B,T,S,heads,L,mode", [(2, 3, 40, 2, 5, 1), (2, 4, 24, 1, 77, 3), (3, 2, 16, 4, 256, 1), (2, 2, 9, 2, 33, 3)])
def test_cross_attn_short_context(ops, B, T, S, heads, L, mode):
    dev = "cuda"
    C, M = heads * 64, B * T * S
    q = torch.randn(M, C, device=dev).to(BF)
    kv = torch.randn(B * L, 2 * C, device=dev).to(BF)
    kw = dict(ctx_mode=1, ctx_div=T * S) if mode == 1 else dict(ctx_mode=3, ctx_div=T * S, ctx_mod=S, ctx_B=B)
    out = ops.cross_attn(q, kv, L, heads, **kw)
    m = torch.arange(M, device=dev)
    ctx = m // (T * S) if mode == 1 else ((m // (T * S)) * S + m % S) % B
    k = kv[:, :C].float().view(B, L, heads, 64)[ctx]          # [M, L, heads, 64]
    v = kv[:, C:].float().view(B, L, heads, 64)[ctx]
    qq = q.float().view(M, heads, 1, 64)
    ref = F.scaled_dot_product_attention(qq, k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3)).reshape(M, C)
    assert not torch.isnan(out).any() and rel(out, ref) < 5e-3


def _gn_table(ops, units, rows, C_total):
    return ops.GNStats(torch.zeros(ops.GNStats.numel(units), dtype=torch.int64, device="cuda"), units, rows, C_total)


def _gn_expected(y, units, rows, C_total, c_off):
    """(sum, sum of squares) per (unit, group) of the STORED bf16 tensor y [units*rows, C] occupying channels
    c_off.. of a GroupNorm(32, C_total) input (float64 on the GPU), and the same sums of magnitudes."""
    cg = C_total // 32
    C = y.shape[1]
    grp = (torch.arange(C, device=y.device) + c_off) // cg
    onehot = torch.zeros(C, 32, dtype=torch.float64, device=y.device)
    onehot[torch.arange(C, device=y.device), grp] = 1.0
    yd = y.double().view(units, rows, C)
    exp = torch.stack([yd.sum(1) @ onehot, (yd * yd).sum(1) @ onehot], -1)  # [units, 32, 2]
    mag = torch.stack([yd.abs().sum(1) @ onehot, (yd * yd).sum(1) @ onehot], -1)
    return exp, mag


def _gn_check(st, expected, n_values):
    """fp32 partial sums of <= 256 values each (relative 2^-22 per add), every partial rounded once to 2^-16."""
    exp, mag = expected
    got = st.total()
    tol = 4e-6 * mag + (n_values / 32.0 + 2.0) * 2.0 ** -16
    assert torch.all((got - exp).abs() <= tol), (float((got - exp).abs().max()), float(tol.min()))


@pytest.mark.parametrize("kind,geom,C,N,C_total,c_off,temporal", [
    ("conv3x3", (4, 16, 16), 64, 320, 320, 0, False),       # 10 channels per group: pieces straddle groups
    ("conv3x3", (28, 5, 8), 64, 128, 128, 0, False),        # 40-row frames: a warp's 32 rows straddle units
    ("conv3x3", (6, 8, 8), 64, 256, 256, 0, True),          # statistics across frames (unit = clip of 3 frames)
    ("conv3x3_s2", (3, 16, 32), 64, 256, 256, 0, False),
    ("linear", (2, 24, 40), 128, 640, 1920, 1280, False),   # skip half of a 1920-wide concat norm (60 per group)
    ("linear_res", (28, 40, 1), 64, 1280, 1280, 0, False),
    ("conv_t3", (2, 14, 160), 64, 192, 192, 0, True),
    ("upconv", (2, 5, 8), 64, 128, 384, 0, False),          # four phase launches add into one table
])
def test_groupnorm_statistics_from_the_producer_epilogue(ops, kind, geom, C, N, C_total, c_off, temporal):
    dev = "cuda"
    F_, H, W = geom
    x = torch.randn(F_ * H * W, C, device=dev).to(BF)
    bias = torch.randn(N, device=dev)
    if kind.startswith("conv3x3"):
        stride = 2 if kind.endswith("s2") else 1
        w = (torch.randn(N, 9 * C, device=dev) / (9 * C) ** 0.5).to(BF)
        rows_frame = (H // stride) * (W // stride)
        run = lambda gn: ops.conv3x3(x, F_, H, W, w, stride=stride, bias=bias, gn=gn)
    elif kind.startswith("linear"):
        w = (torch.randn(N, C, device=dev) / C ** 0.5).to(BF)
        res = torch.randn(F_ * H * W, N, device=dev).to(BF) if kind.endswith("res") else None
        rows_frame = H * W
        run = lambda gn: ops.linear(x, w, bias=bias, res1=res, gn=gn)
    elif kind == "conv_t3":
        w = (torch.randn(N, 3 * C, device=dev) / (3 * C) ** 0.5).to(BF)
        rows_frame = W  # geom = (B, T, HW)
        run = lambda gn: ops.conv_t3(x, F_, H, W, w, bias=bias, gn=gn)
    else:
        wp = ops.pack_upconv3x3(torch.randn(N, C, 3, 3, device=dev) / (9 * C) ** 0.5)
        rows_frame = 4 * H * W
        run = lambda gn: ops.upsample2x_conv3x3(x, F_, H, W, wp, bias=bias, gn=gn)
    if kind == "conv_t3":
        units, rows = F_, H * W
    elif temporal:
        units, rows = F_ // 3, 3 * rows_frame
    else:
        units, rows = F_, rows_frame
    plain = run(None)
    st = _gn_table(ops, units, rows, C_total)
    y = run((st, c_off))
    st2 = _gn_table(ops, units, rows, C_total)
    run((st2, c_off))
    torch.cuda.synchronize()
    assert torch.equal(y, plain)                 # the statistics do not disturb the output
    assert torch.equal(st.buf.view(st.rep, -1).sum(0), st2.buf.view(st.rep, -1).sum(0))  # integer sums: bit-reproducible
    _gn_check(st, _gn_expected(y, units, rows, C_total, c_off), rows * (C_total // 32))


@pytest.mark.parametrize("units,rows,C,C_total,c_off", [(28, 160, 1280, 1920, 640), (3, 2560, 320, 640, 320), (5, 7, 64, 128, 64)])
def test_axpby_with_groupnorm_statistics(ops, units, rows, C, C_total, c_off):
    dev = "cuda"
    x = torch.randn(units * rows, C, device=dev).to(BF)
    r = (torch.randn(units * rows, C, device=dev) * 0.3).to(BF)
    st = _gn_table(ops, units, rows, C_total)
    y = ops.axpby(x, r, gn=(st, c_off))
    torch.cuda.synchronize()
    assert torch.equal(y, ops.axpby(x, r))
    _gn_check(st, _gn_expected(y, units, rows, C_total, c_off), rows * (C_total // 32))


@pytest.mark.parametrize("F_,H,W,C0,C1,silu", [(28, 10, 16, 320, 0, True), (6, 5, 8, 1280, 640, True), (4, 16, 16, 128, 0, False)])
def test_groupnorm_from_producer_statistics_equals_two_pass(ops, F_, H, W, C0, C1, silu):
    """conv -> (residual add) -> GroupNorm with the statistics taken from the producers' epilogues agrees with
    the two-pass GroupNorm of the same stored tensors (and with torch), and is bit-reproducible."""
    dev = "cuda"
    HW, C = H * W, C0 + C1
    xin = torch.randn(F_ * HW, 64, device=dev).to(BF)
    w = (torch.randn(C0, 9 * 64, device=dev) / 24.0).to(BF)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    s_ = torch.randn(F_ * HW, C1, device=dev).to(BF) if C1 else None
    outs = []
    for _ in range(2):
        st = _gn_table(ops, F_, HW, C)
        y0 = ops.conv3x3(xin, F_, H, W, w, gn=(st, 0))
        y1 = ops.axpby(s_, s_, 0.5, 0.25, gn=(st, C0)) if C1 else None
        outs.append((ops.groupnorm(y0, F_, HW, g, b, 1e-6, silu, src1=y1, stats=st), y0, y1))
    torch.cuda.synchronize()
    (fused, y0, y1), (fused2, _, _) = outs
    assert torch.equal(fused, fused2)
    two_pass = ops.groupnorm(y0, F_, HW, g, b, 1e-6, silu, src1=y1)
    assert rel(fused, two_pass) < 2e-3
    x = torch.cat([y0, y1], 1) if C1 else y0
    ref = F.group_norm(x.float().view(F_, HW, C).permute(0, 2, 1), 32, g, b, 1e-6)
    if silu:
        ref = F.silu(ref)
    assert rel(fused, ref.permute(0, 2, 1).reshape(-1, C)) < TOL_BF16


@pytest.mark.parametrize("M,C,rb", [(4096, 320, False), (1120, 1280, False), (14 * 2 * 24, 64, True), (14 * 40, 640, True), (77, 128, False)])
def test_layernorm(ops, M, C, rb):
    dev = "cuda"
    x = (torch.randn(M, C, device=dev) * 1.5 + 0.2).to(BF)
    g = torch.randn(C, device=dev); b = torch.randn(C, device=dev)
    kw, xr = {}, x.float()
    if rb:
        T = 7; S = M // (2 * T)
        pos = torch.randn(T, C, device=dev)
        kw = dict(rowbias=pos, rb_div=S, rb_mod=T)
        xr = xr + pos[(torch.arange(M, device=dev) // S) % T]
    out = ops.layernorm(x, g, b, 1e-5, **kw)
    torch.cuda.synchronize()
    assert rel(out, F.layer_norm(xr, (C,), g, b, 1e-5)) < TOL_BF16


def test_small_linear_and_sinusoid(ops):
    dev = "cuda"
    for M, K, N, ai, ao in [(2, 320, 1280, False, True), (2, 1280, 1280, True, False), (14, 320, 1280, False, True), (25, 1024, 320, False, False)]:
        x = torch.randn(M, K, device=dev); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF); b = torch.randn(N, device=dev)
        ref = (F.silu(x) if ai else x) @ w.float().t() + b
        ref = F.silu(ref) if ao else ref
        out = ops.small_linear(x, w, b, ai, ao)
        assert rel(out, ref) < TOL_F32
        acc = out.clone()
        ops.small_linear(x, w, b, ai, ao, out=acc, accumulate=True)
        assert rel(acc, 2 * ref) < TOL_F32
    t = torch.tensor([1.63777006, -1.55365205, 6.0, 127.0, 0.02], device=dev)
    e = ops.sinusoid(t, 320, round_bf16=False)
    # SURVEY.md A.8 known answer
    assert torch.allclose(e[0, :3].cpu(), torch.tensor([-0.0669237, 0.0246392, 0.1109036]), atol=2e-6)
    assert torch.allclose(e[0, 160:163].cpu(), torch.tensor([0.9977581, 0.9996964, 0.9938312]), atol=2e-6)
    from oracle.svd_oracle import Timesteps
    assert rel(e, Timesteps(320)(t)) < 1e-5


def test_loop_glue(ops):
    from oracle.sampling import EulerDiscreteSchedulerOracle
    dev = "cuda"
    B, T, h, w = 2, 3, 8, 16
    lat = torch.randn(B, T, 4, h, w, device=dev) * 10
    img = torch.randn(2 * B, T, 4, h, w, device=dev); ctl = torch.randn(2 * B, T, 4, h, w, device=dev)
    sch = EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
    i = 9
    sd = sch.sigmas[i:i + 2].to(dev).contiguous()
    out = ops.prep_input(lat, img, ctl, True, sd)
    ref = torch.zeros(2 * B, T, h, w, 64, device=dev)
    ref[..., 0:4] = (torch.cat([lat, lat]) / float((sch.sigmas[i] ** 2 + 1) ** 0.5)).permute(0, 1, 3, 4, 2)
    ref[..., 4:8] = img.permute(0, 1, 3, 4, 2); ref[..., 8:12] = ctl.permute(0, 1, 3, 4, 2)
    assert rel(out, ref.reshape(-1, 64)) < TOL_BF16 and float(out[:, 12:].abs().max()) == 0.0
    noise = torch.randn(2 * B * T * h * w, 4, device=dev)
    g = torch.linspace(1, 3, T, device=dev)
    lat2 = lat.clone()
    ops.cfg_euler(lat2, noise, True, g, sd)
    n5 = noise.view(2 * B, T, h, w, 4).permute(0, 1, 4, 2, 3)
    v = n5[:B] + g.view(1, T, 1, 1, 1) * (n5[B:] - n5[:B])
    sch.step_index = i
    want = sch.step(v, sch.timesteps[i], lat)   # the oracle's EulerDiscreteScheduler.step
    assert rel(lat2, want) < TOL_F32
    x = torch.randn(3 * 4 * 6, 64, device=dev).to(BF)
    up = ops.upsample2x(x, 3, 4, 6)
    refu = F.interpolate(x.float().view(3, 4, 6, 64).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.float(), refu.permute(0, 2, 3, 1).reshape(-1, 64))
    y = torch.randn(3 * 4 * 6, 64, device=dev).to(BF)
    assert rel(ops.axpby(x, y, 0.3, 0.7), 0.3 * x.float() + 0.7 * y.float()) < TOL_BF16


def test_errors_are_loud(ops):
    from ctrlv_b200._lib import CtrlvError
    a = torch.randn(64, 48, device="cuda").to(BF)   # K not a multiple of 64
    w = torch.randn(64, 48, device="cuda").to(BF)
    with pytest.raises(CtrlvError):
        ops.linear(a, w)
    with pytest.raises(ValueError):
        ops.linear(a.float(), w)


@pytest.mark.parametrize("frames,H,W,C,N", [(2, 4, 6, 64, 64), (3, 10, 16, 128, 64), (1, 5, 8, 64, 128), (2, 20, 32, 64, 96)])
def test_upsample2x_conv3x3_fused(ops, frames, H, W, C, N):
    """Upsample2D (nearest 2x + 3x3 conv) as four 2x2 phase convs of the low-res frame."""
    g = torch.Generator("cpu").manual_seed(5)
    x = torch.randn(frames, C, H, W, generator=g).cuda().to(BF)
    w = (torch.randn(N, C, 3, 3, generator=g) / (9 * C) ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    want = F.conv2d(F.interpolate(x.float(), scale_factor=2.0, mode="nearest"), w.to(BF).float(), b, padding=1)
    rows = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    got = ops.upsample2x_conv3x3(rows, frames, H, W, ops.pack_upconv3x3(w), bias=b)
    got = got.float().reshape(frames, 2 * H, 2 * W, N).permute(0, 3, 1, 2)
    assert rel(got, want) < 6e-3, rel(got, want)


# ---- stream-K schedule of the implicit GEMM (ctrlv_epilogue.splitk_ws) -------------------------------------------
def _sk_problem(ops, kind, geom, C, N, extra):
    """-> (run(gn) launching the problem, fp32 reference rows, statistics geometry)"""
    dev = "cuda"
    if kind == "linear":
        M = geom
        a = torch.randn(M, C, device=dev).to(BF)
        w = (torch.randn(N, C, device=dev) / C ** 0.5).to(BF)
        b = torch.randn(N, device=dev)
        r1 = torch.randn(M, N, device=dev).to(BF)
        rb = torch.randn(3, N, device=dev)
        div = max(M // 5, 1)
        ref = 0.6 * (a.float() @ w.float().t() + b + rb[(torch.arange(M, device=dev) // div) % 3]) + 0.7 * r1.float()
        run = lambda gn: ops.linear(a, w, bias=b, res1=r1, s_res1=0.7, s_acc=0.6, rowbias=rb, rb_mode=2, rb_div=div,
                                    rb_mod=3, gn=gn)
        return run, ref, (1, M)
    if kind in ("conv3x3", "conv3x3_s2", "conv3x3_sc"):
        F_, H, W = geom
        stride = 2 if kind.endswith("s2") else 1
        x = torch.randn(F_, C, H, W, device=dev).to(BF)
        w = (torch.randn(N, C, 3, 3, device=dev) / (9 * C) ** 0.5).to(BF)
        b = torch.randn(N, device=dev)
        ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)
        rows = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
        wp, kw = _pack9(w), {}
        if kind.endswith("sc"):  # second source (skip concat) is folded into x here; raw 1x1 shortcut appended to K
            x2 = torch.randn(F_, 64, H, W, device=dev).to(BF)
            ws = (torch.randn(N, 64, device=dev) / 8.0).to(BF)
            ref = ref + F.conv2d(x2.float(), ws.float()[:, :, None, None])
            wp = torch.cat([wp, ws], 1).contiguous()
            kw["sc0"] = rows(x2)
        xl = rows(x)
        run = lambda gn: ops.conv3x3(xl, F_, H, W, wp, stride=stride, bias=b, gn=gn, **kw)
        return run, rows(ref), (F_, (H // stride) * (W // stride))
    B, T, HW = geom
    x = torch.randn(B, C, T, HW, 1, device=dev).to(BF)
    w = (torch.randn(N, C, 3, 1, 1, device=dev) / (3 * C) ** 0.5).to(BF)
    b = torch.randn(N, device=dev)
    res = torch.randn(B * T * HW, N, device=dev).to(BF)
    ref = 0.5 * F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))[..., 0].permute(0, 2, 3, 1).reshape(-1, N) + 0.5 * res.float()
    xl = x[..., 0].permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    wp = w[:, :, :, 0, 0].permute(0, 2, 1).reshape(N, 3 * C).contiguous()
    run = lambda gn: ops.conv_t3(xl, B, T, HW, wp, bias=b, res1=res, s_res1=0.5, s_acc=0.5, gn=gn)
    return run, ref, (B, T * HW)


@pytest.mark.parametrize("kind,geom,C,N", [
    ("linear", 1120, 1280, 1280),            # deepest UNet level: 10 row tiles
    ("linear", 1000, 2048, 320),             # ragged M, 160-wide pair tiles
    ("linear", 100, 1024, 64),               # ONE tile cut between two CTAs
    ("linear", 4096, 1024, 1280),            # ranges spanning tail + head (+ whole tiles) of several tiles
    ("linear", 130, 4096, 96),               # single CTAs would not pair: 2 row tiles, narrow ragged-N tile
    ("conv3x3", (28, 5, 8), 1280, 1280),     # config-2 level-3 conv, K = 11520
    ("conv3x3_sc", (6, 5, 8), 192, 128),     # shortcut segment at the end of the K loop
    ("conv3x3_s2", (4, 10, 16), 256, 256),   # Downsample2D: four parity maps
    ("conv_t3", (2, 14, 40), 1280, 1280),    # config-2 level-3 temporal conv
    ("conv_t3", (1, 5, 24), 448, 64),
])
def test_streamk_schedule_matches_whole_tiles(ops, monkeypatch, kind, geom, C, N):
    """The stream-K schedule (k-ranges per CTA, fp32 partials, fix-up launch) gives the reference result, agrees
    with the whole-tile schedule up to the order of the fp32 partial sums, is bit-reproducible from run to run
    (fixed reduction order, also with a dirty workspace), and accumulates the same GroupNorm statistics in its
    fix-up epilogue."""
    lib = ops.lib()
    monkeypatch.setattr(ops, "SPLITK", True)  # every launch of this test carries a workspace
    run, ref, (units, rows) = _sk_problem(ops, kind, geom, C, N, None)
    C_total, c_off = 2 * N, N  # the output is the upper half of a concatenated norm input
    fus = ops.GNStats.fusable(C_total)
    try:
        assert lib.ctrlv_igemm_streamk(1) == 0
        whole = run(None)
        assert lib.ctrlv_igemm_streamk(2) == 0
        ops._splitk_workspace().fill_(0xFF)  # (NaN patterns: every partial that is read was written by this launch)
        n0 = lib.ctrlv_launch_count()
        y = run(None)
        assert lib.ctrlv_launch_count() == n0 + 2  # GEMM + fix-up
        st = _gn_table(ops, units, rows, C_total) if fus else None
        y2 = run((st, c_off) if fus else None)
        torch.cuda.synchronize()
    finally:
        lib.ctrlv_igemm_streamk(0)
    assert rel(y, ref) < TOL_BF16
    assert rel(y, whole) < 2e-3
    assert torch.equal(y, y2)
    if fus:
        _gn_check(st, _gn_expected(y, units, rows, C_total, c_off), rows * (C_total // 32))


def test_streamk_heuristic_splits_only_badly_filled_long_k_problems(ops, monkeypatch):
    """Without a forced mode only a badly filled problem with a very long K loop is cut along K (the result then
    differs from whole tiles by the order of the partial sums only); everything else keeps its whole-tile schedule
    bit for bit, and so does every problem when no workspace is passed."""
    lib = ops.lib()
    monkeypatch.setattr(ops, "SPLITK", True)
    for (M, K, N), launches in (((1120, 16384, 1280), 2), ((1120, 1280, 1280), 1), ((71680 // 4, 320, 320), 1),
                                ((4480, 5120, 1280), 1)):
        a = torch.randn(M, K, device="cuda").to(BF)
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
        try:
            lib.ctrlv_igemm_streamk(1)
            whole = ops.linear(a, w)
        finally:
            lib.ctrlv_igemm_streamk(0)
        n0 = lib.ctrlv_launch_count()
        auto = ops.linear(a, w)
        assert lib.ctrlv_launch_count() - n0 == launches
        torch.cuda.synchronize()
        assert rel(auto, whole) < 2e-3
        if launches == 1:
            assert torch.equal(auto, whole)
    monkeypatch.setattr(ops, "SPLITK", False)
    a = torch.randn(1120, 16384, device="cuda").to(BF)
    w = (torch.randn(1280, 16384, device="cuda") / 128.0).to(BF)
    n0 = lib.ctrlv_launch_count()
    ops.linear(a, w)
    assert lib.ctrlv_launch_count() - n0 == 1
