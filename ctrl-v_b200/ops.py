"""torch.Tensor-level wrappers over the C ABI (device memory + streams are torch's; the math is
the library's).  Every wrapper launches on torch's current CUDA stream and raises CtrlvError on
failure — there is no eager/PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import Epilogue, check

BF16 = torch.bfloat16


def lib():
    return _lib.load()


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _req(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor")


def make_ep(out=None, out_f32=None, bias=None, rowbias=None, rb_mode=0, rb_div=1, rb_mod=1, rb_B=1,
            geglu=False, s_acc=1.0, res1=None, s_res1=1.0, res2=None, s_res2=1.0, n_store=0) -> Epilogue:
    ep = Epilogue()
    if bias is not None:
        _req(bias, torch.float32, "bias")
    ep.bias = _p(bias)
    if rowbias is not None:
        _req(rowbias, torch.float32, "rowbias")
        ep.ld_rowbias = rowbias.stride(0)
    ep.rowbias = _p(rowbias)
    ep.rb_mode, ep.rb_div, ep.rb_mod, ep.rb_B = rb_mode if rowbias is not None else 0, rb_div, rb_mod, rb_B
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
    check(lib().ctrlv_linear(a.data_ptr(), a.stride(0), M, K, w.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_linear")
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]


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
    check(lib().ctrlv_conv3x3(x.data_ptr(), C0, _p(src1), C1, frames, H, W, stride, _p(sc0), SC0,
                              _p(sc1), SC1, w.data_ptr(), N, C.byref(ep), _stream()), "ctrlv_conv3x3")
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
    check(lib().ctrlv_conv_t3(x.data_ptr(), Cc, B, T, HW, w.data_ptr(), N, C.byref(ep), _stream()),
          "ctrlv_conv_t3")
    return kw["out"] if kw.get("out") is not None else kw["out_f32"]
