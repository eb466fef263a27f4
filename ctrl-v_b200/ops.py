"""torch.Tensor-level wrappers over the C ABI (device memory + streams are torch's; the math is
the library's).  Every wrapper launches on torch's current CUDA stream and raises CtrlvError on
failure — there is no eager/PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import Epilogue, check

BF16 = torch.bfloat16


def lib():
    return _lib.load()


# Optional per-call timing (eager mode only; developer tool): set ops.PROFILE = {} to collect
# {(op, shape-key): [calls, ms, flops]} with CUDA events around every call.
PROFILE = None
_prof_pending = []


def _prof(op, key, flops=0.0, nbytes=0.0):
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _prof_pending.append((op, key, flops, nbytes, e0, e1))
    return e1


def _prof_end(tok):
    if tok is not None:
        tok.record()


def profile_flush():
    torch.cuda.synchronize()
    for op, key, flops, nbytes, e0, e1 in _prof_pending:
        r = PROFILE.setdefault((op, key), [0, 0.0, 0.0, 0.0])  # calls, ms, algorithmic flops, algorithmic bytes
        r[0] += 1; r[1] += e0.elapsed_time(e1); r[2] += flops; r[3] += nbytes
    _prof_pending.clear()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor")


class GNStats:
    """(sum, sum of squares) accumulators of ONE nn.GroupNorm(32, C_total), filled by the launches that
    produce its input (igemm epilogue `gn=`, `axpby(gn=)`) so that the norm runs without a statistics
    pass: int64 fixed point [rep][n_units][32][2] (value * 2^16; integer atomics, hence order-independent and
    bit-reproducible), replicated `rep` times so that the producers' atomics do not all land on the same few
    L2 lines when there are only a few units.  `buf` must be zero before the first producer runs."""

    __slots__ = ("buf", "n_units", "rows_per_unit", "C", "cg", "rep")

    def __init__(self, buf: torch.Tensor, n_units: int, rows_per_unit: int, C_total: int):
        assert buf.dtype == torch.int64 and buf.is_contiguous() and buf.numel() == self.numel(n_units)
        assert C_total % 32 == 0
        self.buf, self.n_units, self.rows_per_unit, self.C, self.cg = buf, n_units, rows_per_unit, C_total, C_total // 32
        self.rep = self.replicas(n_units)

    @staticmethod
    def replicas(n_units: int) -> int:
        """power of two, >= 128 table rows in total"""
        r = 1
        while r * n_units < 128:
            r *= 2
        return r

    @classmethod
    def numel(cls, n_units: int) -> int:
        return cls.replicas(n_units) * n_units * 64

    def total(self) -> torch.Tensor:
        """[n_units, 32, 2] float64 (sum, sum of squares): the replicas added up (test / debugging helper)"""
        return self.buf.view(self.rep, self.n_units, 32, 2).sum(0).double() / 65536.0

    @staticmethod
    def fusable(C_total: int) -> bool:
        """The epilogue splits an aligned 8-column piece between at most two groups (cg = 4, 6 or >= 8) and
        the residual-add kernel reduces channel pairs (even cg)."""
        cg = C_total // 32
        return C_total % 32 == 0 and cg % 2 == 0 and cg >= 4


# Stream-K workspaces of the implicit GEMM (ctrlv_epilogue.splitk_ws): fp32 partial-tile slots, two per CTA; one
# workspace per (device, stream), because launches on different streams may run concurrently.
# Off unless CTRLV_SPLITK=1: on the denoise step's shapes the schedule is measured level with whole tiles (step
# 33.8 vs 33.7 ms, profiles/r02_streamk_microbench.json), so the product path keeps one launch per contraction.
SPLITK = os.environ.get("CTRLV_SPLITK", "0") == "1"
_sk_ws = {}


def _splitk_workspace() -> torch.Tensor:
    dev = torch.cuda.current_device()
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _sk_ws.get(key)
    if ws is None:
        nbytes = 2 * torch.cuda.get_device_properties(dev).multi_processor_count * (128 << 10)
        ws = _sk_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    return ws


def make_ep(out=None, out_f32=None, bias=None, rowbias=None, rb_mode=0, rb_div=1, rb_mod=1, rb_B=1,
            geglu=False, s_acc=1.0, res1=None, s_res1=1.0, res2=None, s_res2=1.0, n_store=0, gn=None,
            rb_off=0) -> Epilogue:
    """gn = (GNStats, c_off): also accumulate the statistics of `out` for the GroupNorm that consumes it,
    `out` being channels c_off.. of that norm's (possibly concatenated) input."""
    ep = Epilogue()
    if gn is not None:
        st, c_off = gn
        ep.gn_sums, ep.gn_rows_per_unit, ep.gn_cg, ep.gn_c_off = st.buf.data_ptr(), st.rows_per_unit, st.cg, c_off
        ep.gn_units, ep.gn_rep = st.n_units, st.rep
    if bias is not None:
        _req(bias, torch.float32, "bias")
    ep.bias = _p(bias)
    if rowbias is not None:
        _req(rowbias, torch.float32, "rowbias")
        ep.ld_rowbias = rowbias.stride(0)
    ep.rowbias = _p(rowbias)
    ep.rb_mode, ep.rb_div, ep.rb_mod, ep.rb_B = rb_mode if rowbias is not None else 0, rb_div, rb_mod, rb_B
    ep.rb_off = rb_off
    ep.geglu = 1 if geglu else 0
    ep.s_acc = s_acc
    for name, t in (("res1", res1), ("res2", res2), ("out", out)):
        if t is not None:
            _req(t, BF16, name)
            if t.stride(-1) != 1:
                raise ValueError(f"{name}: last dim must be contiguous")
    ep.res1, ep.ld_res1, ep.s_res1 = _p(res1), (res1.stride(0) if res1 is not None else 0), s_res1
    ep.res2, ep.ld_res2, ep.s_res2 = _p(res2), (res2.stride(0) if res2 is not None else 0), s_res2
    ep.out, ep.ld_out = _p(out), (out.stride(0) if out is not None else 0)
    if out_f32 is not None:
        _req(out_f32, torch.float32, "out_f32")
    ep.out_f32, ep.ld_out_f32 = _p(out_f32), (out_f32.stride(0) if out_f32 is not None else 0)
    ep.n_store = n_store
    if SPLITK:
        ws = _splitk_workspace()
        ep.splitk_ws, ep.splitk_bytes = ws.data_ptr(), ws.numel()
    return ep


def _alloc_out(M, N, geglu, kw):
    if kw.get("out") is None and kw.get("out_f32") is None:
        kw["out"] = torch.empty((M, N // 2 if geglu else N), dtype=BF16, device="cuda")
    return kw


def linear(a: torch.Tensor, w: torch.Tensor, **kw) -> torch.Tensor:
    """out[M, N] = a[M, K] @ w[N, K]^T with the fused epilogue (see make_ep)."""
    _req(a, BF16, "a"); _req(w, BF16, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous() and a.stride(1) == 1
    kw = _alloc_out(M, N, kw.get("geglu", False), kw)
    ep = make_ep(**kw)
    tok = _prof("linear", (M, K, N, bool(kw.get("geglu")), kw.get("res1") is not None, kw.get("res2") is not None, kw.get("gn") is not None), 2.0 * M * K * N)
    check(lib().ctrlv_linear(a.data_ptr(), a.stride(0), M, K, w.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_linear")
    _prof_end(tok)
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]


FF_FUSED_MAX_C = 320  # ctrlv_feedforward keeps D [128 x C], S [128 x 128] and H [128 x 64] in the 512 TMEM columns


def feedforward(x: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor, ln_eps: Optional[float] = None,
                ln_rowbias: Optional[torch.Tensor] = None, ln_rb_div: int = 1, ln_rb_mod: int = 1, **kw) -> torch.Tensor:
    """out = epilogue(GEGLU(x @ w1^T + b1) @ w2^T) in ONE launch (C <= 320); w1 / b1 with interleaved
    (value, gate) rows; kw = the output epilogue (bias = b2, rowbias, s_acc, res1, res2, out).
    ln_eps: x are the rows BEFORE the LayerNorm in front of the FeedForward (affine part folded into w1 / b1);
    they are normalised tile by tile inside the launch (+ ln_rowbias[(m // ln_rb_div) % ln_rb_mod] first)."""
    _req(x, BF16, "x"); _req(w1, BF16, "w1"); _req(w2, BF16, "w2"); _req(b1, torch.float32, "b1")
    M, Cc = x.shape
    assert x.stride(1) == 1 and w1.is_contiguous() and w2.is_contiguous()
    assert tuple(w1.shape) == (8 * Cc, Cc) and tuple(w2.shape) == (Cc, 4 * Cc) and b1.numel() == 8 * Cc
    kw = _alloc_out(M, Cc, False, kw)
    ep = make_ep(**kw)
    tok = _prof("feedforward", (M, Cc, kw.get("res1") is not None, kw.get("res2") is not None), 2.0 * M * Cc * 12 * Cc)
    if ln_eps is not None:
        if ln_rowbias is not None:
            _req(ln_rowbias, torch.float32, "ln_rowbias")
        check(lib().ctrlv_feedforward_ln(x.data_ptr(), x.stride(0), M, Cc, ln_eps, _p(ln_rowbias),
                                         ln_rowbias.stride(0) if ln_rowbias is not None else 0, ln_rb_div, ln_rb_mod,
                                         w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), C.byref(ep), _stream()),
              "ctrlv_feedforward_ln")
    else:
        check(lib().ctrlv_feedforward(x.data_ptr(), x.stride(0), M, Cc, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                                      C.byref(ep), _stream()), "ctrlv_feedforward")
    _prof_end(tok)
    return kw["out"]


def linear_ln(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, ln_eps: float = 1e-5,
              ln_rowbias: Optional[torch.Tensor] = None, ln_rb_div: int = 1, ln_rb_mod: int = 1,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M, N] = LayerNorm(x [+ ln_rowbias])[M, K] @ w[N, K]^T + bias in ONE launch (K <= 320, N % 64 == 0): the
    rows are normalised tile by tile in shared memory (the norm's affine part folded into w / bias by the caller)."""
    _req(x, BF16, "x"); _req(w, BF16, "w")
    M, K = x.shape
    N = w.shape[0]
    assert w.shape[1] == K and w.is_contiguous() and x.stride(1) == 1 and K <= FF_FUSED_MAX_C and N % 64 == 0
    if out is None:
        out = torch.empty((M, N), dtype=BF16, device="cuda")
    ep = make_ep(out=out, bias=bias)
    if ln_rowbias is not None:
        _req(ln_rowbias, torch.float32, "ln_rowbias")
    tok = _prof("linear", (M, K, N, False, False, False, False), 2.0 * M * K * N)
    check(lib().ctrlv_linear_ln(x.data_ptr(), x.stride(0), M, K, ln_eps, _p(ln_rowbias),
                                ln_rowbias.stride(0) if ln_rowbias is not None else 0, ln_rb_div, ln_rb_mod,
                                w.data_ptr(), N, C.byref(ep), _stream()), "ctrlv_linear_ln")
    _prof_end(tok)
    return out


def conv3x3(x: torch.Tensor, frames: int, H: int, W: int, w: torch.Tensor, stride: int = 1,
            src1: Optional[torch.Tensor] = None, sc0: Optional[torch.Tensor] = None,
            sc1: Optional[torch.Tensor] = None, **kw) -> torch.Tensor:
    """3x3 / pad 1 conv on channels-last rows x[frames*H*W, C0] (optionally | src1), weights
    w[N, 9*(C0+C1) + SC0 + SC1] tap-major; sc0/sc1 are raw inputs of a fused 1x1 shortcut."""
    _req(x, BF16, "x"); _req(w, BF16, "w")
    assert x.is_contiguous() and w.is_contiguous()
    C0 = x.shape[1]
    C1 = src1.shape[1] if src1 is not None else 0
    SC0 = sc0.shape[1] if sc0 is not None else 0
    SC1 = sc1.shape[1] if sc1 is not None else 0
    for t in (src1, sc0, sc1):
        if t is not None:
            _req(t, BF16, "src"); assert t.is_contiguous()
    N = w.shape[0]
    assert w.shape[1] == 9 * (C0 + C1) + SC0 + SC1, (w.shape, C0, C1, SC0, SC1)
    Mo = frames * (H // stride) * (W // stride)
    kw = _alloc_out(Mo, N, kw.get("geglu", False), kw)
    ep = make_ep(**kw)
    tok = _prof("conv3x3", (frames, H, W, stride, C0 + C1, SC0 + SC1, N, kw.get("gn") is not None), 2.0 * Mo * w.shape[1] * N)
    check(lib().ctrlv_conv3x3(x.data_ptr(), C0, _p(src1), C1, frames, H, W, stride, _p(sc0), SC0,
                              _p(sc1), SC1, w.data_ptr(), N, C.byref(ep), _stream()), "ctrlv_conv3x3")
    _prof_end(tok)
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]


def conv_t3(x: torch.Tensor, B: int, T: int, HW: int, w: torch.Tensor, **kw) -> torch.Tensor:
    """(3,1,1) temporal conv, zero halo per clip, on x[B*T*HW, C]; weights w[N, 3*C] tap-major."""
    _req(x, BF16, "x"); _req(w, BF16, "w")
    assert x.is_contiguous() and w.is_contiguous()
    Cc = x.shape[1]
    N = w.shape[0]
    assert w.shape[1] == 3 * Cc and x.shape[0] == B * T * HW
    kw = _alloc_out(B * T * HW, N, kw.get("geglu", False), kw)
    ep = make_ep(**kw)
    tok = _prof("conv_t3", (B, T, HW, Cc, N, kw.get("gn") is not None), 2.0 * B * T * HW * 3 * Cc * N)
    check(lib().ctrlv_conv_t3(x.data_ptr(), Cc, B, T, HW, w.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_conv_t3")
    _prof_end(tok)
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]


_gn_ws = {}


def _gn_workspace(n_units: int) -> torch.Tensor:
    key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    need = lib().ctrlv_groupnorm_workspace(n_units)
    ws = _gn_ws.get(key)
    if ws is None or ws.numel() * 4 < need:
        ws = torch.empty(max(need // 4, 1 << 16), dtype=torch.float32, device="cuda")
        _gn_ws[key] = ws
    return ws


def groupnorm(x: torch.Tensor, n_units: int, rows_per_unit: int, gamma: torch.Tensor, beta: torch.Tensor,
              eps: float, silu: bool, src1: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None,
              stats: Optional[GNStats] = None) -> torch.Tensor:
    """GroupNorm(32) [+SiLU] over statistics units of `rows_per_unit` rows; x (| src1) -> out.
    `stats`: the statistics were accumulated by the producers of x (| src1): normalise only."""
    _req(x, BF16, "x"); _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    assert x.is_contiguous() and x.shape[0] == n_units * rows_per_unit
    C0 = x.shape[1]
    C1 = 0
    if src1 is not None:
        _req(src1, BF16, "src1"); assert src1.is_contiguous() and src1.shape[0] == x.shape[0]
        C1 = src1.shape[1]
    assert gamma.numel() == C0 + C1 and beta.numel() == C0 + C1
    if out is None:
        out = torch.empty((x.shape[0], C0 + C1), dtype=BF16, device="cuda")
    if stats is not None:
        assert (stats.n_units, stats.rows_per_unit, stats.C) == (n_units, rows_per_unit, C0 + C1), "GNStats mismatch"
        tok = _prof("groupnorm_apply", (n_units, rows_per_unit, C0 + C1), nbytes=4.0 * x.shape[0] * (C0 + C1))
        check(lib().ctrlv_groupnorm_apply(x.data_ptr(), C0, _p(src1), C1, n_units, rows_per_unit, gamma.data_ptr(),
                                          beta.data_ptr(), eps, 1 if silu else 0, out.data_ptr(),
                                          stats.buf.data_ptr(), stats.rep, _stream()), "ctrlv_groupnorm_apply")
        _prof_end(tok)
        return out
    if ws is None:
        ws = _gn_workspace(n_units)
    tok = _prof("groupnorm", (n_units, rows_per_unit, C0 + C1), nbytes=4.0 * x.shape[0] * (C0 + C1))
    check(lib().ctrlv_groupnorm(x.data_ptr(), C0, _p(src1), C1, n_units, rows_per_unit, gamma.data_ptr(),
                                beta.data_ptr(), eps, 1 if silu else 0, out.data_ptr(), ws.data_ptr(),
                                _stream()), "ctrlv_groupnorm")
    _prof_end(tok)
    return out


def layernorm(x: torch.Tensor, gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
              eps: float = 1e-5, rowbias: Optional[torch.Tensor] = None, rb_div: int = 1, rb_mod: int = 1,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm over rows; gamma/beta None = normalisation only (affine folded into the consumer)."""
    _req(x, BF16, "x")
    if gamma is not None:
        _req(gamma, torch.float32, "gamma"); _req(beta, torch.float32, "beta")
    M, Cc = x.shape
    assert x.stride(1) == 1
    if out is None:
        out = torch.empty((M, Cc), dtype=BF16, device="cuda")
    if rowbias is not None:
        _req(rowbias, torch.float32, "rowbias")
    tok = _prof("layernorm", (M, Cc), nbytes=4.0 * M * Cc)
    check(lib().ctrlv_layernorm(x.data_ptr(), x.stride(0), M, Cc, _p(gamma), _p(beta), eps,
                                _p(rowbias), rowbias.stride(0) if rowbias is not None else 0, rb_div,
                                rb_mod, out.data_ptr(), _stream()), "ctrlv_layernorm")
    _prof_end(tok)
    return out


def attn_spatial(qkv: torch.Tensor, frames: int, S: int, heads: int, scale: Optional[float] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(qkv, BF16, "qkv")
    Cc = heads * 64
    assert qkv.is_contiguous() and qkv.shape == (frames * S, 3 * Cc), (qkv.shape, frames, S, heads)
    if out is None:
        out = torch.empty((frames * S, Cc), dtype=BF16, device="cuda")
    tok = _prof("attn_spatial", (frames, S, heads), 4.0 * frames * heads * S * S * 64, nbytes=8.0 * frames * S * Cc)
    check(lib().ctrlv_attn_spatial(qkv.data_ptr(), frames, S, heads, scale if scale is not None else 0.125,
                                   out.data_ptr(), _stream()), "ctrlv_attn_spatial")
    _prof_end(tok)
    return out


def attn_temporal(qkv: torch.Tensor, B: int, T: int, S: int, heads: int, scale: Optional[float] = None,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(qkv, BF16, "qkv")
    Cc = heads * 64
    assert qkv.is_contiguous() and qkv.shape == (B * T * S, 3 * Cc)
    if out is None:
        out = torch.empty((B * T * S, Cc), dtype=BF16, device="cuda")
    tok = _prof("attn_temporal", (B, T, S, heads), 4.0 * B * S * heads * T * T * 64, nbytes=8.0 * B * T * S * Cc)
    check(lib().ctrlv_attn_temporal(qkv.data_ptr(), B, T, S, heads, scale if scale is not None else 0.125,
                                    out.data_ptr(), _stream()), "ctrlv_attn_temporal")
    _prof_end(tok)
    return out


def cross_attn(q: torch.Tensor, kv: torch.Tensor, L: int, heads: int, ctx_mode: int = 1, ctx_div: int = 1,
               ctx_mod: int = 1, ctx_B: int = 1, scale: Optional[float] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Cross-attention over L > 1 context tokens: q [M, C] (projected), kv [n_ctx*L, 2C] = to_k | to_v of the
    contexts; row m attends to context ridx(m) (make_ep's rb_mode conventions)."""
    _req(q, BF16, "q"); _req(kv, BF16, "kv")
    M, Cc = q.shape
    assert Cc == heads * 64 and kv.shape[1] == 2 * Cc and kv.shape[0] % L == 0 and q.stride(1) == 1 and kv.stride(1) == 1
    if out is None:
        out = torch.empty((M, Cc), dtype=BF16, device="cuda")
    check(lib().ctrlv_cross_attn(q.data_ptr(), q.stride(0), kv.data_ptr(), kv.stride(0), M, heads, L, kv.shape[0] // L,
                                 scale if scale is not None else 0.125, ctx_mode, ctx_div, ctx_mod, ctx_B,
                                 out.data_ptr(), _stream()), "ctrlv_cross_attn")
    return out


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 act_in: bool = False, act_out: bool = False, out: Optional[torch.Tensor] = None,
                 accumulate: bool = False) -> torch.Tensor:
    """y (+)= act_out(act_in(x) @ w^T + b) for M <= 64 rows; x, y fp32, w bf16."""
    _req(x, torch.float32, "x"); _req(w, BF16, "w")
    M, K = x.shape
    N = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and w.shape[1] == K
    if out is None:
        assert not accumulate
        out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    check(lib().ctrlv_small_linear(x.data_ptr(), M, K, w.data_ptr(), _p(bias), N, 1 if act_in else 0,
                                   1 if act_out else 0, 1 if accumulate else 0, out.data_ptr(), _stream()),
          "ctrlv_small_linear")
    return out


def sinusoid(t: torch.Tensor, dim: int, round_bf16: bool = True) -> torch.Tensor:
    _req(t, torch.float32, "t")
    t = t.contiguous()
    out = torch.empty((t.numel(), dim), dtype=torch.float32, device="cuda")
    check(lib().ctrlv_sinusoid(t.data_ptr(), t.numel(), dim, 1 if round_bf16 else 0, out.data_ptr(), _stream()),
          "ctrlv_sinusoid")
    return out


def prep_input(latents: torch.Tensor, image_latents: Optional[torch.Tensor], control_cond: Optional[torch.Tensor],
               cfg: bool, sigma_dev: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sigma_dev: fp32 CUDA tensor, element 0 = sigma of this step."""
    _req(latents, torch.float32, "latents"); _req(sigma_dev, torch.float32, "sigma_dev")
    B, T, c4, h, w = latents.shape
    assert c4 == 4 and latents.is_contiguous()
    nb = 2 * B if cfg else B
    for t in (image_latents, control_cond):
        if t is not None:
            _req(t, torch.float32, "cond"); assert t.is_contiguous() and t.shape == (nb, T, 4, h, w), t.shape
    if out is None:
        out = torch.empty((nb * T * h * w, 64), dtype=BF16, device="cuda")
    check(lib().ctrlv_prep_input(latents.data_ptr(), _p(image_latents), _p(control_cond), B, 1 if cfg else 0,
                                 T, h, w, sigma_dev.data_ptr(), out.data_ptr(), _stream()), "ctrlv_prep_input")
    return out


def cfg_euler(latents: torch.Tensor, noise: torch.Tensor, cfg: bool, guidance: Optional[torch.Tensor],
              sigma_dev: torch.Tensor, round_bf16: bool = False) -> torch.Tensor:
    """sigma_dev: fp32 CUDA tensor [sigma, sigma_next]; latents updated in place."""
    _req(latents, torch.float32, "latents"); _req(noise, torch.float32, "noise")
    _req(sigma_dev, torch.float32, "sigma_dev"); assert sigma_dev.numel() >= 2
    B, T, c4, h, w = latents.shape
    assert latents.is_contiguous() and noise.stride(1) == 1
    check(lib().ctrlv_cfg_euler(latents.data_ptr(), noise.data_ptr(), noise.stride(0), B, 1 if cfg else 0, T, h,
                                w, _p(guidance), sigma_dev.data_ptr(), 1 if round_bf16 else 0,
                                _stream()), "ctrlv_cfg_euler")
    return latents


def upsample2x(x: torch.Tensor, frames: int, H: int, W: int) -> torch.Tensor:
    _req(x, BF16, "x")
    Cc = x.shape[1]
    assert x.is_contiguous() and x.shape[0] == frames * H * W
    out = torch.empty((frames * 4 * H * W, Cc), dtype=BF16, device="cuda")
    check(lib().ctrlv_upsample2x(x.data_ptr(), frames, H, W, Cc, out.data_ptr(), _stream()), "ctrlv_upsample2x")
    return out


def axpby(x: torch.Tensor, y: torch.Tensor, a: float = 1.0, b: float = 1.0,
          out: Optional[torch.Tensor] = None, gn=None) -> torch.Tensor:
    """out = a*x + b*y (bf16); gn = (GNStats, c_off) also accumulates the statistics of `out` (rows [M, C])."""
    _req(x, BF16, "x"); _req(y, BF16, "y")
    assert x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()
    if out is None:
        out = torch.empty_like(x)
    if gn is not None:
        st, c_off = gn
        M, Cc = x.shape
        assert M == st.n_units * st.rows_per_unit
        check(lib().ctrlv_axpby_gn(x.data_ptr(), y.data_ptr(), a, b, M, Cc, out.data_ptr(), st.buf.data_ptr(),
                                   st.rows_per_unit, st.cg, c_off, st.rep, _stream()), "ctrlv_axpby_gn")
        return out
    check(lib().ctrlv_axpby(x.data_ptr(), y.data_ptr(), a, b, x.numel(), out.data_ptr(), _stream()), "ctrlv_axpby")
    return out


def memset_zero(t: torch.Tensor) -> torch.Tensor:
    """Zero a contiguous tensor through the library (recordable by a launch plan, unlike torch's fill kernel)."""
    assert t.is_contiguous() and t.is_cuda
    check(lib().ctrlv_memset_zero(t.data_ptr(), t.numel() * t.element_size(), _stream()), "ctrlv_memset_zero")
    return t


def gn_stats_of(x: torch.Tensor, gn) -> None:
    """Accumulate the GroupNorm statistics of x [M, C] itself into gn = (GNStats, c_off) — the same reduction
    as `axpby(gn=)`, for a tensor whose producer could not (a skip connection with two consumers)."""
    _req(x, BF16, "x")
    assert x.is_contiguous()
    st, c_off = gn
    M, Cc = x.shape
    assert M == st.n_units * st.rows_per_unit
    check(lib().ctrlv_axpby_gn(x.data_ptr(), None, 1.0, 0.0, M, Cc, None, st.buf.data_ptr(), st.rows_per_unit, st.cg,
                               c_off, st.rep, _stream()), "ctrlv_axpby_gn")


def nchw_to_nhwc(src: torch.Tensor, out: torch.Tensor, c_off: int = 0) -> torch.Tensor:
    """src [frames, C, H, W] (fp32/bf16, contiguous) -> columns c_off.. of out [frames*H*W, Cpad] bf16."""
    assert src.is_contiguous() and src.dtype in (torch.float32, BF16)
    frames, Cc, H, W = src.shape
    _req(out, BF16, "out")
    assert out.is_contiguous() and out.shape[0] == frames * H * W
    check(lib().ctrlv_nchw_to_nhwc(src.data_ptr(), 1 if src.dtype == torch.float32 else 0, frames, Cc, H * W,
                                   out.shape[1], c_off, out.data_ptr(), _stream()), "ctrlv_nchw_to_nhwc")
    return out


def nhwc_to_nchw(src: torch.Tensor, frames: int, Cc: int, H: int, W: int, dtype=torch.float32) -> torch.Tensor:
    assert src.stride(1) == 1 and src.dtype in (torch.float32, BF16) and dtype in (torch.float32, BF16)
    out = torch.empty((frames, Cc, H, W), dtype=dtype, device="cuda")
    check(lib().ctrlv_nhwc_to_nchw(src.data_ptr(), 1 if src.dtype == torch.float32 else 0, src.stride(0), frames,
                                   Cc, H * W, 1 if dtype == torch.float32 else 0, out.data_ptr(), _stream()),
          "ctrlv_nhwc_to_nchw")
    return out


# ---- temporal VAE glue ---------------------------------------------------------
def conv3x3_s2_pad01(x: torch.Tensor, frames: int, H: int, W: int, w: torch.Tensor, **kw) -> torch.Tensor:
    """diffusers `Downsample2D(padding=0)` of the VAE encoder: F.pad(x, (0,1,0,1)) then a 3x3 stride-2
    conv without padding, i.e. out(oy,ox) = sum_k w[ky][kx] x(2oy+ky, 2ox+kx).  Described to the
    generic implicit GEMM as four parity sub-lattices: tap k -> parity k&1, lattice offset k>>1; the
    out-of-range row/column of the last tap is the TMA zero fill."""
    _req(x, BF16, "x"); _req(w, BF16, "w")
    assert x.is_contiguous() and w.is_contiguous() and H % 2 == 0 and W % 2 == 0
    C0 = x.shape[1]
    N = w.shape[0]
    assert C0 % 64 == 0 and w.shape[1] == 9 * C0 and x.shape[0] == frames * H * W
    kw = _alloc_out(frames * (H // 2) * (W // 2), N, False, kw)
    ep = make_ep(**kw)
    check(lib().ctrlv_conv3x3_s2_pad01(x.data_ptr(), C0, frames, H, W, w.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_conv3x3_s2_pad01")
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]


def softmax_rows(scores: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(scores, torch.float32, "scores")
    M, N = scores.shape
    assert scores.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), dtype=BF16, device="cuda")
    check(lib().ctrlv_softmax_rows(scores.data_ptr(), scores.stride(0), M, N, scale, out.data_ptr(), out.stride(0),
                                   _stream()), "ctrlv_softmax_rows")
    return out


def time_conv_out(x: torch.Tensor, B: int, T: int, H: int, W: int, Cc: int, w: torch.Tensor,
                  bias: torch.Tensor) -> torch.Tensor:
    """x [B*T*H*W, ld] fp32 channels-last -> [B*T, Cc, H, W] fp32 after Conv3d(Cc, Cc, (3,1,1))."""
    _req(x, torch.float32, "x"); _req(w, torch.float32, "w"); _req(bias, torch.float32, "bias")
    assert x.stride(1) == 1 and x.shape[0] == B * T * H * W and w.is_contiguous() and w.numel() == Cc * Cc * 3
    out = torch.empty((B * T, Cc, H, W), dtype=torch.float32, device="cuda")
    check(lib().ctrlv_time_conv_out(x.data_ptr(), x.stride(0), B, T, H * W, Cc, w.data_ptr(), bias.data_ptr(),
                                    out.data_ptr(), _stream()), "ctrlv_time_conv_out")
    return out


# ---- image-conditioning prologue ------------------------------------------------
def blur1d_reflect(x: torch.Tensor, taps: torch.Tensor, axis: int) -> torch.Tensor:
    """x [..., H, W] fp32; one pass of the separable Gaussian (reflect padding) along x (0) or y (1)."""
    _req(x, torch.float32, "x"); _req(taps, torch.float32, "taps")
    assert x.is_contiguous() and taps.is_contiguous()
    H, W = x.shape[-2:]
    out = torch.empty_like(x)
    check(lib().ctrlv_blur1d_reflect(x.data_ptr(), x.numel() // (H * W), H, W, axis, taps.data_ptr(), taps.numel(),
                                     out.data_ptr(), _stream()), "ctrlv_blur1d_reflect")
    return out


def resize_bicubic_ac(x: torch.Tensor, Ho: int, Wo: int) -> torch.Tensor:
    _req(x, torch.float32, "x")
    assert x.is_contiguous()
    H, W = x.shape[-2:]
    out = torch.empty(x.shape[:-2] + (Ho, Wo), dtype=torch.float32, device="cuda")
    check(lib().ctrlv_resize_bicubic_ac(x.data_ptr(), x.numel() // (H * W), H, W, Ho, Wo, out.data_ptr(), _stream()),
          "ctrlv_resize_bicubic_ac")
    return out


def clip_patchify(img: torch.Tensor, patch: int, mean: torch.Tensor, std: torch.Tensor, a: float = 1.0,
                  s: float = 0.0, clamp01: bool = False) -> torch.Tensor:
    """img [B, C, H, W] fp32 -> bf16 rows [B*(H/P)*(W/P), Kpad] (Kpad = C*P*P rounded up to 64)."""
    _req(img, torch.float32, "img"); _req(mean, torch.float32, "mean"); _req(std, torch.float32, "std")
    assert img.is_contiguous()
    B, Cc, H, W = img.shape
    Kpad = (Cc * patch * patch + 63) // 64 * 64
    out = torch.empty((B * (H // patch) * (W // patch), Kpad), dtype=BF16, device="cuda")
    check(lib().ctrlv_clip_patchify(img.data_ptr(), B, Cc, H, W, patch, a, s, 1 if clamp01 else 0, mean.data_ptr(),
                                    std.data_ptr(), Kpad, out.data_ptr(), _stream()), "ctrlv_clip_patchify")
    return out


# ---- fused nearest-2x upsample + 3x3 conv (diffusers Upsample2D) --------------------------------------
def pack_upconv3x3(w: torch.Tensor, device="cuda", dtype=BF16) -> torch.Tensor:
    """Conv2d weight [N, C, 3, 3] of `Upsample2D.conv` -> four phase matrices [4, N, 4*C] (bf16).

    conv3x3(nearest2x(x)) at output pixel (2y+py, 2x+px) only ever sees a 2x2 patch of x: for py = 0 the
    taps ky = 0 | 1, 2 fall on rows y-1 | y, for py = 1 the taps ky = 0, 1 | 2 fall on rows y | y+1 (same
    along x; the zero padding of the upsampled frame coincides with that of x).  Each phase is therefore a
    2x2 conv of the LOW-resolution frame with summed taps: 2.25x fewer MACs and no upsampled intermediate."""
    w = w.detach().float()
    N, Cc = w.shape[0], w.shape[1]
    comb = {0: [(0,), (1, 2)], 1: [(0, 1), (2,)]}  # phase -> two (tap group) sums, in increasing source offset
    out = torch.empty((4, N, 4 * Cc), dtype=torch.float32, device=w.device)
    for py in (0, 1):
        for px in (0, 1):
            blocks = []
            for ky in comb[py]:
                for kx in comb[px]:
                    acc = torch.zeros((N, Cc), dtype=torch.float32, device=w.device)
                    for a in ky:
                        for b in kx:
                            acc += w[:, :, a, b]
                    blocks.append(acc)
            out[py * 2 + px] = torch.cat(blocks, dim=1)
    return out.to(dtype=dtype).contiguous().to(device)


def upsample2x_conv3x3(x: torch.Tensor, frames: int, H: int, W: int, wp: torch.Tensor,
                       bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                       gn=None) -> torch.Tensor:
    """x [frames*H*W, C] -> conv3x3(nearest2x(x)) as rows [frames*2H*2W, N]; wp from pack_upconv3x3."""
    _req(x, BF16, "x"); _req(wp, BF16, "wp")
    assert x.is_contiguous() and wp.is_contiguous() and x.shape[0] == frames * H * W
    C0 = x.shape[1]
    N = wp.shape[1]
    assert C0 % 64 == 0 and tuple(wp.shape) == (4, N, 4 * C0), (wp.shape, C0)
    if out is None:
        out = torch.empty((frames * 4 * H * W, N), dtype=BF16, device="cuda")
    tok = _prof("upconv3x3", (frames, H, W, C0, N), 2.0 * frames * H * W * 16 * C0 * N)
    ep = make_ep(out=out, bias=bias, gn=gn)
    check(lib().ctrlv_upsample2x_conv3x3(x.data_ptr(), C0, frames, H, W, wp.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_upsample2x_conv3x3")
    _prof_end(tok)
    return out
