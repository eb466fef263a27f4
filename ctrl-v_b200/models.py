"""Drop-in models of the Box2Video hot path on the sm_100a kernels.

Mirrors the reference's module interface for this path (same class names, forward signatures,
argument meaning and error behaviour):

  ControlNetModel.forward(sample, timestep, encoder_hidden_states, added_time_ids, control_cond,
                          conditioning_scale, return_dict)
        <- /root/reference/src/ctrlv/models/controlnet.py:226-351
  UNetSpatioTemporalConditionModel.forward(sample, timestep, encoder_hidden_states,
                          added_time_ids, down_block_additional_residuals,
                          mid_block_additional_residuals, return_dict)
        <- /root/reference/src/ctrlv/models/unet_spatio_temporal_condition.py:31-171

Weights live in a flat dict keyed by the diffusers state-dict names and are
repacked once into kernel layouts (bf16 K-major matrices, fp32 biases / norm affine).  Every
contraction, norm and attention below is a call into libctrlv_b200.so — there is no PyTorch
math on the path (torch provides device memory, views and streams only).

Activations are channels-last: a reference tensor [B*T, C, h, w] is a [B*T*h*w, C] bf16 matrix.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

BF16 = torch.bfloat16

SVD_CONFIG = dict(
    sample_size=None, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                      "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                    "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20), num_frames=25,
)


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


# ------------------------------------------------------------------------------------------------
# parameter specification (names + shapes of the diffusers state dict)
# ------------------------------------------------------------------------------------------------
def _spec_linear(spec, name, i, o, bias=True):
    spec[name + ".weight"] = (o, i)
    if bias:
        spec[name + ".bias"] = (o,)


def _spec_norm(spec, name, c):
    spec[name + ".weight"] = (c,)
    spec[name + ".bias"] = (c,)


def _spec_resblock(spec, pfx, cin, cout, temb):
    s = pfx + ".spatial_res_block"
    _spec_norm(spec, s + ".norm1", cin)
    spec[s + ".conv1.weight"] = (cout, cin, 3, 3); spec[s + ".conv1.bias"] = (cout,)
    _spec_linear(spec, s + ".time_emb_proj", temb, cout)
    _spec_norm(spec, s + ".norm2", cout)
    spec[s + ".conv2.weight"] = (cout, cout, 3, 3); spec[s + ".conv2.bias"] = (cout,)
    if cin != cout:
        spec[s + ".conv_shortcut.weight"] = (cout, cin, 1, 1); spec[s + ".conv_shortcut.bias"] = (cout,)
    t = pfx + ".temporal_res_block"
    _spec_norm(spec, t + ".norm1", cout)
    spec[t + ".conv1.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv1.bias"] = (cout,)
    _spec_linear(spec, t + ".time_emb_proj", temb, cout)
    _spec_norm(spec, t + ".norm2", cout)
    spec[t + ".conv2.weight"] = (cout, cout, 3, 1, 1); spec[t + ".conv2.bias"] = (cout,)
    spec[pfx + ".time_mixer.mix_factor"] = (1,)


def _spec_attn(spec, pfx, dim, kv_dim):
    _spec_linear(spec, pfx + ".to_q", dim, dim, bias=False)
    _spec_linear(spec, pfx + ".to_k", kv_dim, dim, bias=False)
    _spec_linear(spec, pfx + ".to_v", kv_dim, dim, bias=False)
    _spec_linear(spec, pfx + ".to_out.0", dim, dim)


def _spec_ff(spec, pfx, dim):
    _spec_linear(spec, pfx + ".net.0.proj", dim, 8 * dim)
    _spec_linear(spec, pfx + ".net.2", 4 * dim, dim)


def _spec_transformer(spec, pfx, c, xdim):
    _spec_norm(spec, pfx + ".norm", c)
    _spec_linear(spec, pfx + ".proj_in", c, c)
    b = pfx + ".transformer_blocks.0"
    _spec_norm(spec, b + ".norm1", c); _spec_attn(spec, b + ".attn1", c, c)
    _spec_norm(spec, b + ".norm2", c); _spec_attn(spec, b + ".attn2", c, xdim)
    _spec_norm(spec, b + ".norm3", c); _spec_ff(spec, b + ".ff", c)
    t = pfx + ".temporal_transformer_blocks.0"
    _spec_norm(spec, t + ".norm_in", c); _spec_ff(spec, t + ".ff_in", c)
    _spec_norm(spec, t + ".norm1", c); _spec_attn(spec, t + ".attn1", c, c)
    _spec_norm(spec, t + ".norm2", c); _spec_attn(spec, t + ".attn2", c, xdim)
    _spec_norm(spec, t + ".norm3", c); _spec_ff(spec, t + ".ff", c)
    _spec_linear(spec, pfx + ".time_pos_embed.linear_1", c, 4 * c)
    _spec_linear(spec, pfx + ".time_pos_embed.linear_2", 4 * c, c)
    spec[pfx + ".time_mixer.mix_factor"] = (1,)
    _spec_linear(spec, pfx + ".proj_out", c, c)


def param_spec(cfg: dict, controlnet: bool) -> "OrderedDict[str, tuple]":
    """Names and shapes of the diffusers-format state dict of the UNet (or ControlNet)."""
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    heads = _tup(cfg["num_attention_heads"], n)
    xdim = _tup(cfg["cross_attention_dim"], n)
    lpb = _tup(cfg["layers_per_block"], n)
    temb = boc[0] * 4
    spec["conv_in.weight"] = (boc[0], cfg["in_channels"], 3, 3); spec["conv_in.bias"] = (boc[0],)
    _spec_linear(spec, "time_embedding.linear_1", boc[0], temb)
    _spec_linear(spec, "time_embedding.linear_2", temb, temb)
    _spec_linear(spec, "add_embedding.linear_1", cfg["projection_class_embeddings_input_dim"], temb)
    _spec_linear(spec, "add_embedding.linear_2", temb, temb)
    out_ch = boc[0]
    for i, t in enumerate(cfg["down_block_types"]):
        in_ch, out_ch = out_ch, boc[i]
        for j in range(lpb[i]):
            _spec_resblock(spec, f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch, out_ch, temb)
        if t.startswith("CrossAttn"):
            for j in range(lpb[i]):
                _spec_transformer(spec, f"down_blocks.{i}.attentions.{j}", out_ch, xdim[i])
        if i != n - 1:
            spec[f"down_blocks.{i}.downsamplers.0.conv.weight"] = (out_ch, out_ch, 3, 3)
            spec[f"down_blocks.{i}.downsamplers.0.conv.bias"] = (out_ch,)
    _spec_resblock(spec, "mid_block.resnets.0", boc[-1], boc[-1], temb)
    _spec_transformer(spec, "mid_block.attentions.0", boc[-1], xdim[-1])
    _spec_resblock(spec, "mid_block.resnets.1", boc[-1], boc[-1], temb)
    if controlnet:
        spec["control_conv_in.weight"] = (boc[0], cfg["in_channels"] // 2, 3, 3)
        spec["control_conv_in.bias"] = (boc[0],)
        chans = [boc[0]]
        for i, ch in enumerate(boc):
            chans += [ch] * lpb[i]
            if i != n - 1:
                chans.append(ch)
        for k, ch in enumerate(chans):
            spec[f"controlnet_down_blocks.{k}.weight"] = (ch, ch, 1, 1)
            spec[f"controlnet_down_blocks.{k}.bias"] = (ch,)
        spec["controlnet_mid_block.weight"] = (boc[-1], boc[-1], 1, 1)
        spec["controlnet_mid_block.bias"] = (boc[-1],)
        return spec
    rboc, rheads, rlpb, rxdim = boc[::-1], heads[::-1], lpb[::-1], xdim[::-1]
    out_ch = rboc[0]
    for i, t in enumerate(cfg["up_block_types"]):
        prev, out_ch = out_ch, rboc[i]
        in_ch = rboc[min(i + 1, n - 1)]
        nl = rlpb[i] + 1
        for j in range(nl):
            skip = in_ch if j == nl - 1 else out_ch
            rin = prev if j == 0 else out_ch
            _spec_resblock(spec, f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch, temb)
        if t.startswith("CrossAttn"):
            for j in range(nl):
                _spec_transformer(spec, f"up_blocks.{i}.attentions.{j}", out_ch, rxdim[i])
        if i != n - 1:
            spec[f"up_blocks.{i}.upsamplers.0.conv.weight"] = (out_ch, out_ch, 3, 3)
            spec[f"up_blocks.{i}.upsamplers.0.conv.bias"] = (out_ch,)
    _spec_norm(spec, "conv_norm_out", boc[0])
    spec["conv_out.weight"] = (cfg["out_channels"], boc[0], 3, 3); spec["conv_out.bias"] = (cfg["out_channels"],)
    return spec


def random_state_dict(cfg: dict, controlnet: bool, seed: int = 0, device="cuda",
                      dtype=BF16, zero_conv_std: float = 0.0) -> Dict[str, torch.Tensor]:
    """Random-init weights with PyTorch's default layer statistics (uniform(+-1/sqrt(fan_in)),
    norm affine = (1, 0), mix_factor = 0.5), drawn directly on the device.  The ControlNet
    zero-convs are zeros like the reference's `zero_module` (controlnet.py:149-184), so a fresh or
    `from_unet` ControlNet is a no-op on the UNet; benchmarks and parity code pass
    `zero_conv_std=0.02` so that the injection path carries signal."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd = {}
    spec = param_spec(cfg, controlnet)
    for name, shape in spec.items():
        if name.endswith("mix_factor"):
            t = torch.full(shape, 0.5, device=device)
        elif len(shape if name.endswith("weight") else spec[name[:-4] + "weight"]) == 1:  # norm affine
            t = torch.ones(shape, device=device) if name.endswith("weight") else torch.zeros(shape, device=device)
        elif name.startswith("controlnet_"):
            t = (torch.randn(shape, device=device, generator=g) * zero_conv_std if zero_conv_std
                 else torch.zeros(shape, device=device))
        else:
            wshape = shape if name.endswith("weight") else spec[name[:-4] + "weight"]
            fan_in = 1
            for d in wshape[1:]:
                fan_in *= d
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, device=device, generator=g) * 2 - 1) * bound
        sd[name] = t.to(dtype)
    return sd


# ------------------------------------------------------------------------------------------------
# weight packing
# ------------------------------------------------------------------------------------------------
def _w(t):  # matrix operand (converted where it lives, then one plain copy to the GPU)
    return t.detach().to(dtype=BF16).contiguous().to("cuda")


def _f(t):  # fp32 vector
    return t.detach().to(dtype=torch.float32).contiguous().to("cuda")


def _conv9(w):  # [Cout, Cin, 3, 3] -> [Cout, 9*Cin], tap-major
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _conv_t3(w):  # [Cout, Cin, 3, 1, 1] -> [Cout, 3*Cin]
    return w[:, :, :, 0, 0].permute(0, 2, 1).reshape(w.shape[0], -1)


def _interleave_geglu(w, b):
    """GEGLU proj rows [0:inner] are values, [inner:2*inner] gates (diffusers `chunk(2, -1)`);
    interleave them so a (value, gate) pair sits in adjacent GEMM columns."""
    inner = w.shape[0] // 2
    wi = torch.stack([w[:inner], w[inner:]], dim=1).reshape(w.shape[0], w.shape[1])
    bi = torch.stack([b[:inner], b[inner:]], dim=1).reshape(-1)
    return wi, bi


class _Lin:
    def __init__(self, sd, name, bias=True):
        self.w = _w(sd[name + ".weight"].reshape(sd[name + ".weight"].shape[0], -1))
        self.b = _f(sd[name + ".bias"]) if bias else None


class _Norm:
    def __init__(self, sd, name):
        self.g = _f(sd[name + ".weight"])
        self.b = _f(sd[name + ".bias"])


class _ResBlock:
    def __init__(self, sd, pfx, eps):
        s, t = pfx + ".spatial_res_block", pfx + ".temporal_res_block"
        self.eps = eps
        self.norm1, self.norm2 = _Norm(sd, s + ".norm1"), _Norm(sd, s + ".norm2")
        w1 = sd[s + ".conv1.weight"]
        self.cin, self.cout = w1.shape[1], w1.shape[0]
        self.conv1_w, self.conv1_b = _w(_conv9(w1.float())), _f(sd[s + ".conv1.bias"])
        self.temb = _Lin(sd, s + ".time_emb_proj")
        w2 = _conv9(sd[s + ".conv2.weight"].float())
        b2 = sd[s + ".conv2.bias"].float()
        self.has_shortcut = (s + ".conv_shortcut.weight") in sd
        if self.has_shortcut:  # 1x1 shortcut rides in the K loop of conv2
            ws = sd[s + ".conv_shortcut.weight"].float().reshape(self.cout, self.cin)
            w2 = torch.cat([w2, ws.to(w2.device)], dim=1)
            b2 = b2 + sd[s + ".conv_shortcut.bias"].float().to(b2.device)
        self.conv2_w, self.conv2_b = _w(w2), _f(b2)
        self.tnorm1, self.tnorm2 = _Norm(sd, t + ".norm1"), _Norm(sd, t + ".norm2")
        self.tconv1_w, self.tconv1_b = _w(_conv_t3(sd[t + ".conv1.weight"].float())), _f(sd[t + ".conv1.bias"])
        self.ttemb = _Lin(sd, t + ".time_emb_proj")
        self.tconv2_w, self.tconv2_b = _w(_conv_t3(sd[t + ".conv2.weight"].float())), _f(sd[t + ".conv2.bias"])
        self.alpha = float(torch.sigmoid(sd[pfx + ".time_mixer.mix_factor"].float()).item())

    def __call__(self, x, aux, g, x1=None, st_in=None, st_out=None):
        """x [M, C0] (| x1 [M, C1] skip connection) -> [M, Cout];  g = geometry (B, T, H, W);
        aux.temb holds every block's time_emb_proj(silu(emb)) (one batched launch per forward).

        GroupNorm statistics ride in the epilogue of the launch that PRODUCES a norm's input
        (SURVEY.md §8 row g1): `st_in` = statistics table of norm1's input x (| x1) if the caller's
        producers accumulated it; `st_out` = table of the GroupNorm that consumes this block's output
        (None: no norm follows, or not fused), filled by the last conv here.  Each of the four norms
        below therefore runs as one apply pass."""
        B, T, H, W = g
        F_, HW = B * T, H * W
        a1 = ops.groupnorm(x, F_, HW, self.norm1.g, self.norm1.b, self.eps, True, src1=x1, stats=st_in)
        temb = aux.temb[:, self.temb_off:self.temb_off + self.cout]
        ttemb = aux.temb[:, self.ttemb_off:self.ttemb_off + self.cout]
        st = aux.gn_stats(F_, HW, self.cout)
        h = ops.conv3x3(a1, F_, H, W, self.conv1_w, bias=self.conv1_b, rowbias=temb, rb_mode=1,
                        rb_div=T * HW, gn=_gn(st))
        a2 = ops.groupnorm(h, F_, HW, self.norm2.g, self.norm2.b, self.eps, True, stats=st)
        st = aux.gn_stats(B, T * HW, self.cout)
        if self.has_shortcut:
            xs = ops.conv3x3(a2, F_, H, W, self.conv2_w, sc0=x, sc1=x1, bias=self.conv2_b, gn=_gn(st))
        else:
            assert x1 is None
            xs = ops.conv3x3(a2, F_, H, W, self.conv2_w, bias=self.conv2_b, res1=x, gn=_gn(st))
        # temporal resnet on [B, T, HW, C] (GroupNorm statistics across frames), then AlphaBlender:
        #   out = a*xs + (1-a)*(xs + h_t) = xs + (1-a)*h_t
        a3 = ops.groupnorm(xs, B, T * HW, self.tnorm1.g, self.tnorm1.b, self.eps, True, stats=st)
        st = aux.gn_stats(B, T * HW, self.cout)
        h2 = ops.conv_t3(a3, B, T, HW, self.tconv1_w, bias=self.tconv1_b, rowbias=ttemb, rb_mode=1,
                         rb_div=T * HW, gn=_gn(st))
        a4 = ops.groupnorm(h2, B, T * HW, self.tnorm2.g, self.tnorm2.b, self.eps, True, stats=st)
        return ops.conv_t3(a4, B, T, HW, self.tconv2_w, bias=self.tconv2_b, s_acc=1.0 - self.alpha,
                           res1=xs, s_res1=1.0, gn=_gn(st_out))


def _gn(st, c_off: int = 0):
    """epilogue argument `gn=` for a producer of channels c_off.. of the norm whose statistics `st` holds"""
    return None if st is None else (st, c_off)


# GroupNorm statistics from the producers' epilogues (True) or from gn_stats_kernel passes (False: the
# round-1 path, kept for A/B measurements and as the parity cross-check of the fused statistics)
GN_FUSED = os.environ.get("CTRLV_GN_FUSED", "1") != "0"
# FeedForward as one fused launch where the width allows it (True) or as two igemm launches (False: A/B runs)
FF_FUSED = os.environ.get("CTRLV_FF_FUSED", "1") != "0"
FF_LN = os.environ.get("CTRLV_FF_LN", "1") != "0"  # the LayerNorm in front of a fused FeedForward runs inside its launch
QKV_LN = os.environ.get("CTRLV_QKV_LN", "1") != "0"  # norm1 + the fused q | k | v projection in one launch (C <= 320)


def _qkv(h, attn):
    """LayerNorm(norm1, affine part folded) + the fused to_q | to_k | to_v projection of a self-attention"""
    if QKV_LN and h.shape[1] <= ops.FF_FUSED_MAX_C and attn.wqkv.shape[0] % 64 == 0:
        return ops.linear_ln(h, attn.wqkv, bias=attn.bqkv)
    return ops.linear(ops.layernorm(h), attn.wqkv, bias=attn.bqkv)
if os.environ.get("CTRLV_FF_CG"):  # A/B runs: force single CTAs (1) or CTA pairs (2) in the fused FeedForward
    from . import _lib as _l
    _l.check(_l.load().ctrlv_feedforward_override(int(os.environ["CTRLV_FF_CG"])))


class _Aux:
    """Per-forward side inputs of the blocks: the batched time-embedding projections and 1-token context
    vectors, plus the arena the GroupNorm statistics tables of this forward come from."""

    def __init__(self, temb, ctx, ctx_all=None, vB=None, b0=0, arena=None, L=1, ehs_rows=None):
        self.temb, self.ctx = temb, ctx
        self.L, self.ehs_rows = L, ehs_rows  # L > 1: the contexts as bf16 rows [n_ctx*L, D] for real cross-attention
        self.ctx_all = ctx if ctx_all is None else ctx_all
        self.vB = ctx.shape[0] if vB is None else vB  # contexts in the whole (virtual) batch
        self.b0 = b0
        self.arena = arena

    def gn_stats(self, n_units: int, rows_per_unit: int, C_total):
        """A zeroed statistics table for the GroupNorm(32, C_total) over `n_units` units (None: not fused)."""
        return None if self.arena is None else self.arena.take(n_units, rows_per_unit, C_total)

    def gn_reset(self):
        """Re-zero the arena and hand its tables out again (callers that reuse one _Aux for several forwards)."""
        if self.arena is not None and self.arena.enabled:
            ops.memset_zero(self.arena.buf)
            self.arena.off = 0


class _GNArena:
    """Zeroed int64 storage for the GroupNorm statistics of one forward (ops.GNStats tables, one per
    norm), handed out in call order so that a captured CUDA graph sees the same addresses on every
    replay; ONE fill kernel per forward zeroes all of it."""

    def __init__(self, n_tables: int, max_units: int, enabled: bool = True):
        self.enabled = enabled
        per_table = max(ops.GNStats.numel(u) for u in range(1, max_units + 1)) if enabled else 0
        self.buf = (ops.memset_zero(torch.empty((n_tables * per_table,), dtype=torch.int64, device="cuda"))
                    if enabled else None)
        self.off = 0

    def take(self, n_units: int, rows_per_unit: int, C_total):
        if not self.enabled or C_total is None or not ops.GNStats.fusable(C_total):
            return None
        n = ops.GNStats.numel(n_units)
        if self.off + n > self.buf.numel():
            raise RuntimeError("GroupNorm statistics arena exhausted")
        st = ops.GNStats(self.buf[self.off:self.off + n], n_units, rows_per_unit, C_total)
        self.off += n
        return st


def _fold_ln(w, b, norm_w, norm_b):
    """Linear(LayerNorm_affine(z)) = (W diag(gamma)) z + (b + W beta): fold the LayerNorm's affine
    part into the Linear that consumes it, so the norm kernel only normalises."""
    w = w.float()
    g, be = norm_w.float().to(w.device), norm_b.float().to(w.device)
    b2 = w @ be
    if b is not None:
        b2 = b2 + b.float().to(w.device)
    return w * g[None, :], b2


class _FF:
    def __init__(self, sd, pfx, norm=None):
        w1, b1 = sd[pfx + ".net.0.proj.weight"].float(), sd[pfx + ".net.0.proj.bias"].float()
        if norm is not None:
            w1, b1 = _fold_ln(w1, b1, sd[norm + ".weight"], sd[norm + ".bias"])
        w1, b1 = _interleave_geglu(w1, b1)
        self.w1, self.b1 = _w(w1), _f(b1)
        self.w2, self.b2 = _w(sd[pfx + ".net.2.weight"]), _f(sd[pfx + ".net.2.bias"])

    def __call__(self, x, ln=None, **kw):
        """[LayerNorm +] GEGLU up-projection + down-projection with the output epilogue `kw` (res1, rowbias, ...):
        one fused launch where the width allows it (the rows are normalised tile by tile in shared memory, the
        4C-wide intermediate stays in tensor memory), otherwise a LayerNorm launch and two igemm launches.
        ln = None: x is already normalised; ln = dict of ops.layernorm's row-bias arguments (possibly empty): x are
        the rows BEFORE the norm whose affine part this FeedForward has folded into its weights."""
        fused = FF_FUSED and x.shape[1] <= ops.FF_FUSED_MAX_C
        if ln is not None and not (fused and FF_LN):
            x, ln = ops.layernorm(x, **ln), None
        if fused:
            if ln is not None:
                return ops.feedforward(x, self.w1, self.b1, self.w2, bias=self.b2, ln_eps=1e-5,
                                       ln_rowbias=ln.get("rowbias"), ln_rb_div=ln.get("rb_div", 1),
                                       ln_rb_mod=ln.get("rb_mod", 1), **kw)
            return ops.feedforward(x, self.w1, self.b1, self.w2, bias=self.b2, **kw)
        return ops.linear(ops.linear(x, self.w1, bias=self.b1, geglu=True), self.w2, bias=self.b2, **kw)


class _SelfAttn:
    def __init__(self, sd, pfx, norm):
        w = torch.cat([sd[pfx + ".to_q.weight"], sd[pfx + ".to_k.weight"], sd[pfx + ".to_v.weight"]], 0)
        w, b = _fold_ln(w, None, sd[norm + ".weight"], sd[norm + ".bias"])
        self.wqkv, self.bqkv = _w(w), _f(b)
        self.out = _Lin(sd, pfx + ".to_out.0")


class _CrossAttnL1:
    """Cross-attention over a 1-token context: softmax over one key is exactly 1, so
    attn2(x, ctx) = to_out(to_v(ctx)) — a per-sample vector."""

    def __init__(self, sd, pfx):
        # to_out(to_v(ctx)) = (W_out W_v) ctx + b_out: fold the two projections (fp32 product)
        self.w = sd[pfx + ".to_out.0.weight"].float() @ sd[pfx + ".to_v.weight"].float()  # [C, xdim]
        self.b = sd[pfx + ".to_out.0.bias"].float()
        self.C = self.w.shape[0]
        self.off = None  # column offset inside the model-wide batched context table


class _CrossAttn:
    """General cross-attention (context of L > 1 tokens): to_q with its LayerNorm's affine part folded in,
    to_k | to_v fused into one projection of the context, to_out.  Packed lazily — the Box2Video pipelines only
    ever pass a 1-token context, which _CrossAttnL1 handles without any of this."""

    def __init__(self, sd, pfx, norm):
        wq, bq = _fold_ln(sd[pfx + ".to_q.weight"], None, sd[norm + ".weight"], sd[norm + ".bias"])
        self.wq, self.bq = _w(wq), _f(bq)
        self.wkv = _w(torch.cat([sd[pfx + ".to_k.weight"], sd[pfx + ".to_v.weight"]], 0))
        self.out = _Lin(sd, pfx + ".to_out.0")

    def __call__(self, h, aux, heads, **ctx_index):
        """h + to_out(softmax(to_q(LN(h)) to_k(ctx)^T / 8) to_v(ctx)); ctx_index: which context a row attends to"""
        q = ops.linear(ops.layernorm(h), self.wq, bias=self.bq)
        kv = ops.linear(aux.ehs_rows, self.wkv)
        o = ops.cross_attn(q, kv, aux.L, heads, **ctx_index)
        return ops.linear(o, self.out.w, bias=self.out.b, res1=h)


class _Transformer:
    def __init__(self, sd, pfx, heads, time_context_order):
        self.heads = heads
        self.order = time_context_order
        self._sd, self._pfx = sd, pfx  # (the owner's state dict, not a copy: general cross-attention packs lazily)
        self._xattn = None
        self.norm = _Norm(sd, pfx + ".norm")
        self.proj_in, self.proj_out = _Lin(sd, pfx + ".proj_in"), _Lin(sd, pfx + ".proj_out")
        b = pfx + ".transformer_blocks.0"
        # every LayerNorm feeds a Linear: its affine part is folded into that Linear's weights
        self.attn1 = _SelfAttn(sd, b + ".attn1", b + ".norm1")
        self.attn2, self.ff = _CrossAttnL1(sd, b + ".attn2"), _FF(sd, b + ".ff", b + ".norm3")
        t = pfx + ".temporal_transformer_blocks.0"
        self.tff_in, self.tff = _FF(sd, t + ".ff_in", t + ".norm_in"), _FF(sd, t + ".ff", t + ".norm3")
        self.tattn1, self.tattn2 = _SelfAttn(sd, t + ".attn1", t + ".norm1"), _CrossAttnL1(sd, t + ".attn2")
        self.pos1, self.pos2 = _Lin(sd, pfx + ".time_pos_embed.linear_1"), _Lin(sd, pfx + ".time_pos_embed.linear_2")
        self.alpha = float(torch.sigmoid(sd[pfx + ".time_mixer.mix_factor"].float()).item())
        self.C = self.proj_in.w.shape[0]
        self._pos_cache: Dict[int, torch.Tensor] = {}

    def xattn(self):
        if self._xattn is None:
            b, t = self._pfx + ".transformer_blocks.0", self._pfx + ".temporal_transformer_blocks.0"
            self._xattn = (_CrossAttn(self._sd, b + ".attn2", b + ".norm2"), _CrossAttn(self._sd, t + ".attn2", t + ".norm2"))
        return self._xattn

    def pos_emb(self, T):
        """time_pos_embed(time_proj(arange(T))): depends on weights and T only."""
        if T not in self._pos_cache:
            idx = torch.arange(T, device="cuda", dtype=torch.float32)
            e = ops.sinusoid(idx, self.C, round_bf16=True)
            self._pos_cache[T] = ops.small_linear(ops.small_linear(e, self.pos1.w, self.pos1.b, act_out=True),
                                                  self.pos2.w, self.pos2.b)
        return self._pos_cache[T]

    def __call__(self, x, aux, g, st_in=None, st_out=None):
        """x [M, C] rows ordered (b, t, site); aux.ctx holds every block's 1-token cross-attention
        vector to_out(to_v(ehs[b])) (one batched launch per forward).  st_in / st_out: GroupNorm statistics
        tables of x (from its producer) / of the norm that consumes the result (see _ResBlock.__call__)."""
        B, T, H, W = g
        S, F_ = H * W, B * T
        a = ops.groupnorm(x, F_, S, self.norm.g, self.norm.b, 1e-6, False, stats=st_in)
        h = ops.linear(a, self.proj_in.w, bias=self.proj_in.b)
        # --- BasicTransformerBlock (spatial)
        qkv = _qkv(h, self.attn1)
        att = ops.attn_spatial(qkv, F_, S, self.heads)
        general = aux.L > 1  # context of several tokens: real cross-attention; one token: a per-sample vector
        if general:
            h = ops.linear(att, self.attn1.out.w, bias=self.attn1.out.b, res1=h)
            h = self.xattn()[0](h, aux, self.heads, ctx_mode=1, ctx_div=T * S)
        else:
            ctx = aux.ctx[:, self.attn2.off:self.attn2.off + self.C]
            h = ops.linear(att, self.attn1.out.w, bias=self.attn1.out.b, rowbias=ctx, rb_mode=1, rb_div=T * S, res1=h)
        h = self.ff(h, ln={}, res1=h)
        # --- TemporalBasicTransformerBlock on h + pos[t]; sequences are the T frames of a site
        pos = self.pos_emb(T)
        hm = self.tff_in(h, ln=dict(rowbias=pos, rb_div=S, rb_mod=T), res1=h, rowbias=pos, rb_mode=2, rb_div=S, rb_mod=T)
        qkv = _qkv(hm, self.tattn1)
        att = ops.attn_temporal(qkv, B, T, S, self.heads)
        if general:
            hm = ops.linear(att, self.tattn1.out.w, bias=self.tattn1.out.b, res1=hm)
            if self.order == "s_major":  # context of row (b, s) is that of clip (b*S + s) % B (see below)
                hm = self.xattn()[1](hm, aux, self.heads, ctx_mode=3, ctx_div=T * S, ctx_mod=S, ctx_B=aux.vB)
            else:
                hm = self.xattn()[1](hm, aux, self.heads, ctx_mode=1, ctx_div=T * S)
        elif self.order == "s_major":  # diffusers 0.27.2: context of row (b, s) is ctx[(b*S + s) % B]
            # over the WHOLE batch (vB rows); a branch-sharded process owns global rows b0 .. b0+B-1, so
            # its local index (b*S + s) is offset by b0*S
            # (an index offset, `rb_off`: no rotated copy of the table)
            ctx_t = aux.ctx_all[:, self.tattn2.off:self.tattn2.off + self.C]
            kw = dict(rb_mode=3, rb_div=T * S, rb_mod=S, rb_B=aux.vB, rb_off=(aux.b0 * S) % aux.vB)
        else:
            ctx_t = aux.ctx[:, self.tattn2.off:self.tattn2.off + self.C]
            kw = dict(rb_mode=1, rb_div=T * S)
        if not general:
            hm = ops.linear(att, self.tattn1.out.w, bias=self.tattn1.out.b, rowbias=ctx_t, res1=hm, **kw)
        # ff(norm3(hm)) + hm, then AlphaBlender: a*h + (1-a)*(ff + hm)
        h = self.tff(hm, ln={}, s_acc=1.0 - self.alpha, res1=hm, s_res1=1.0 - self.alpha, res2=h, s_res2=self.alpha)
        return ops.linear(h, self.proj_out.w, bias=self.proj_out.b, res1=x, gn=_gn(st_out))


class _TimeEmbed:
    def __init__(self, sd, cfg):
        self.t1, self.t2 = _Lin(sd, "time_embedding.linear_1"), _Lin(sd, "time_embedding.linear_2")
        self.a1, self.a2 = _Lin(sd, "add_embedding.linear_1"), _Lin(sd, "add_embedding.linear_2")
        self.dim0 = cfg["block_out_channels"][0]
        self.add_dim = cfg["addition_time_embed_dim"]

    def __call__(self, timesteps, added_time_ids):  # [B] fp32, [B, 3] fp32 -> [B, 4*C0] fp32
        Bn = timesteps.shape[0]
        te = ops.sinusoid(timesteps, self.dim0, round_bf16=True)
        emb = ops.small_linear(ops.small_linear(te, self.t1.w, self.t1.b, act_out=True), self.t2.w, self.t2.b)
        ae = ops.sinusoid(added_time_ids.reshape(-1), self.add_dim, round_bf16=True).reshape(Bn, -1)
        # emb = time_embedding(t_emb) + add_embedding(time_embeds)   (controlnet.py:277-283)
        return ops.small_linear(ops.small_linear(ae, self.a1.w, self.a1.b, act_out=True), self.a2.w, self.a2.b,
                                out=emb, accumulate=True)


class _Output(SimpleNamespace):
    pass


class _PackedModel(torch.nn.Module):
    """Common part: config, flat diffusers-format state dict, packing, embeddings, encoder."""

    is_controlnet = False

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None,
                 time_context_order: str = "s_major", seed: int = 0, zero_conv_std: float = 0.0, **overrides):
        super().__init__()
        cfg = dict(SVD_CONFIG)
        if self.is_controlnet:
            cfg.pop("out_channels"); cfg.pop("up_block_types")
        cfg.update(overrides)
        boc = cfg["block_out_channels"]
        if len(boc) != len(cfg["down_block_types"]):  # controlnet.py:80-83
            raise ValueError(
                f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                f"`block_out_channels`: {boc}. `down_block_types`: {cfg['down_block_types']}.")
        if not isinstance(cfg["num_attention_heads"], int) and len(cfg["num_attention_heads"]) != len(boc):
            raise ValueError("Must provide the same number of `num_attention_heads` as `down_block_types`.")
        for c in boc:
            if c % 64 != 0:
                raise ValueError(f"block_out_channels must be multiples of 64 for the sm_100a kernels, got {boc}")
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.time_context_order = time_context_order
        self.dtype = BF16
        self._sd: Dict[str, torch.Tensor] = {}
        self._version = 0  # bumped by every (re)pack: cached CUDA graphs hold raw pointers of the packed weights
        if state_dict is None:
            state_dict = random_state_dict(cfg, self.is_controlnet, seed=seed, zero_conv_std=zero_conv_std)
        self.load_state_dict(state_dict)
        # `add_embedding.linear_1.in_features` is read by ctrlv.utils.util:161
        self.add_embedding = SimpleNamespace(linear_1=SimpleNamespace(
            in_features=cfg["projection_class_embeddings_input_dim"]))

    # ---- state dict with diffusers key names -------------------------------------------------
    def state_dict(self, *a, **k):
        return OrderedDict((n, v.to("cuda")) for n, v in self._sd.items())

    def load_state_dict(self, sd, strict: bool = True):
        spec = param_spec(self.cfg, self.is_controlnet)
        missing = [k for k in spec if k not in sd]
        unexpected = [k for k in sd if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        big = sum(int(v.numel()) for v in sd.values()) > 100_000_000
        for k, shape in spec.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
                # small models are packed where their tensors live (permutes, concatenations, LayerNorm folds
                # on the host, only the packed matrices are copied over); big ones are packed on the GPU
                self._sd[k] = sd[k].detach().to("cuda") if big else sd[k].detach()
        self._pack()
        self._version += 1
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    # ---- diffusers-format checkpoints --------------------------------------
    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None,
                        variant: Optional[str] = None, time_context_order: str = "s_major", **kwargs):
        """Load `<path>/<subfolder>/{config.json, diffusion_pytorch_model[.variant].safetensors|.bin}`
        like `ModelMixin.from_pretrained` does for the reference
        (tools/eval_video_controlnet.py:114-118).  Local directories only (no hub)."""
        from . import checkpoint
        kwargs.pop("torch_dtype", None); kwargs.pop("low_cpu_mem_usage", None)
        config, sd = checkpoint.load_diffusers_dir(pretrained_model_name_or_path, subfolder, variant)
        known = set(SVD_CONFIG) - ({"out_channels", "up_block_types"} if cls.is_controlnet else set())
        over = {k: (tuple(v) if isinstance(v, list) else v) for k, v in config.items() if k in known}
        over.update(kwargs)
        return cls(state_dict=sd, time_context_order=time_context_order, **over)

    def save_pretrained(self, save_directory: str, subfolder: Optional[str] = None,
                        variant: Optional[str] = None, safe_serialization: bool = True,
                        max_shard_bytes: Optional[int] = None):
        from . import checkpoint
        name = "ControlNetModel" if self.is_controlnet else "UNetSpatioTemporalConditionModel"
        checkpoint.save_diffusers_dir(save_directory, self.cfg, self._sd, name, subfolder, variant,
                                      safe_serialization, max_shard_bytes)

    def to(self, *a, **k):  # weights are device-resident bf16 by construction
        return self

    def eval(self):
        return self

    @property
    def device(self):
        return torch.device("cuda", torch.cuda.current_device())

    # ---- packing ------------------------------------------------------------------------------
    def _pack_conv_in(self):
        sd, cfg = self._sd, self.cfg
        c0 = cfg["block_out_channels"][0]
        cin = cfg["in_channels"]
        w = torch.zeros(c0, 3, 3, 64, device=sd["conv_in.weight"].device, dtype=torch.float32)
        w[..., :cin] = sd["conv_in.weight"].float().permute(0, 2, 3, 1)
        b = sd["conv_in.bias"].float()
        if self.is_controlnet:  # conv_in(sample) + control_conv_in(cond) == one conv on [sample | cond]
            cc = sd["control_conv_in.weight"].shape[1]
            w[..., cin:cin + cc] = sd["control_conv_in.weight"].float().permute(0, 2, 3, 1)
            b = b + sd["control_conv_in.bias"].float()
        self.conv_in_w, self.conv_in_b = _w(w.reshape(c0, -1)), _f(b)

    def _pack(self):
        sd, cfg = self._sd, self.cfg
        boc = tuple(cfg["block_out_channels"])
        n = len(boc)
        heads = _tup(cfg["num_attention_heads"], n)
        lpb = _tup(cfg["layers_per_block"], n)
        if _tup(cfg["transformer_layers_per_block"], n) != (1,) * n:
            raise NotImplementedError("transformer_layers_per_block != 1 is not supported")
        order = self.time_context_order
        self._pack_conv_in()
        self.embed = _TimeEmbed(sd, cfg)
        self.down = []
        for i, t in enumerate(cfg["down_block_types"]):
            attn = t.startswith("CrossAttn")
            eps = 1e-6 if attn else 1e-5
            res = [_ResBlock(sd, f"down_blocks.{i}.resnets.{j}", eps) for j in range(lpb[i])]
            att = [_Transformer(sd, f"down_blocks.{i}.attentions.{j}", heads[i], order) for j in range(lpb[i])] if attn else None
            ds = _Lin(sd, f"down_blocks.{i}.downsamplers.0.conv") if i != n - 1 else None
            if ds is not None:
                ds.w = _w(_conv9(sd[f"down_blocks.{i}.downsamplers.0.conv.weight"].float()))
            self.down.append((res, att, ds))
        self.mid = (_ResBlock(sd, "mid_block.resnets.0", 1e-5),
                    _Transformer(sd, "mid_block.attentions.0", heads[-1], order),
                    _ResBlock(sd, "mid_block.resnets.1", 1e-5))

    # ---- shared forward pieces -----------------------------------------------------------------
    @staticmethod
    def _check_inputs(sample, timestep, encoder_hidden_states, added_time_ids):
        if not torch.is_tensor(timestep):
            # the reference dropped diffusers' float/int branch and calls `timesteps.shape`
            # (controlnet.py:262-264) -> AttributeError there; be explicit here
            raise TypeError("`timestep` must be a torch.Tensor (0-d or [batch])")
        if sample.dim() != 5:
            raise ValueError(f"`sample` must be [batch, frames, channels, height, width], got {tuple(sample.shape)}")
        if encoder_hidden_states.dim() != 3:
            raise ValueError(f"`encoder_hidden_states` must be [batch, tokens, cross_attention_dim], got "
                             f"{tuple(encoder_hidden_states.shape)}")
        if encoder_hidden_states.shape[1] > 256:
            raise NotImplementedError("contexts longer than 256 tokens are not supported by ctrlv_cross_attn")
        h, w = sample.shape[-2:]
        if h % 8 != 0 or w % 8 != 0:
            raise ValueError(f"latent height and width have to be divisible by 8 but are {h} and {w}.")

    def _embed(self, sample, timestep, added_time_ids):
        Bn = sample.shape[0]
        ts = timestep
        if ts.dim() == 0:
            ts = ts[None]
        ts = ts.to(device="cuda", dtype=torch.float32).expand(Bn).contiguous()
        ids = added_time_ids.to(device="cuda", dtype=torch.float32).contiguous()
        return self.embed(ts, ids)  # [B, 4*C0] fp32

    def _aux(self, emb, ehs, branch: Optional[int] = None, n_units: int = 0):
        """All per-sample vectors of one forward in two batched launches:
        temb[b] = every time_emb_proj(silu(emb[b])); ctx[b] = every to_out(to_v(ehs[b])).

        `branch` = 0/1 selects the CFG-branch-sharded mode: this process runs only the
        uncond (0) or cond (1) half of the CFG batch — `emb` has the local B rows — while `ehs` still
        holds the contexts of BOTH halves [2B, D], because the diffusers-0.27.2 `time_context` order
        pairs row (b, s) with context (b*S + s) % 2B of the whole batch (Appendix A.5)."""
        temb = ops.small_linear(emb, self.temb_w, self.temb_b, act_in=True)
        arena = _GNArena(self._n_gn, n_units, enabled=GN_FUSED) if n_units else None
        if ehs.dim() == 3 and ehs.shape[1] > 1:
            # a context of several tokens (controlnet.py:230,244-245 allow it; the pipelines never produce it):
            # real cross-attention in every transformer block, contexts kept as bf16 rows
            if branch is not None:
                raise NotImplementedError("CFG-branch-sharded forward with a multi-token context")
            Bc, L, D = ehs.shape
            return _Aux(temb, None, ctx_all=temb, vB=Bc, arena=arena, L=L,
                        ehs_rows=ehs.reshape(Bc * L, D).to(BF16).contiguous())
        ehs = ehs.reshape(ehs.shape[0], -1)
        ctx = ops.small_linear(ehs, self.ctx_w, self.ctx_b)
        if branch is None:
            return _Aux(temb, ctx, arena=arena)
        Bl = emb.shape[0]
        if ehs.shape[0] != 2 * Bl:
            raise ValueError(f"branch-sharded forward needs the contexts of both CFG halves: got {ehs.shape[0]} rows "
                             f"for a local batch of {Bl}")
        return _Aux(temb, ctx[branch * Bl:(branch + 1) * Bl], ctx_all=ctx, vB=2 * Bl, b0=branch * Bl, arena=arena)

    def _finish_pack(self, resblocks, transformers):
        tw, tb, off = [], [], 0
        for r in resblocks:
            for lin, name in ((r.temb, "temb_off"), (r.ttemb, "ttemb_off")):
                setattr(r, name, off)
                tw.append(lin.w); tb.append(lin.b); off += lin.w.shape[0]
            r.temb = r.ttemb = None
        self.temb_w, self.temb_b = torch.cat(tw).contiguous(), torch.cat(tb).contiguous()
        cw, cb, off = [], [], 0
        for t in transformers:
            for ca in (t.attn2, t.tattn2):
                ca.off = off
                cw.append(ca.w); cb.append(ca.b); off += ca.C
                ca.w = ca.b = None
        self.ctx_w, self.ctx_b = _w(torch.cat(cw)), _f(torch.cat(cb))
        # statistics tables one forward can ask for: four per ResBlock (norm2, the two temporal norms, the
        # consumer of its output), one per transformer output, conv_in / down- / upsamplers, the skip adds
        self._n_gn = 4 * len(resblocks) + len(transformers) + 40

    def _conv_in(self, inp64, aux, g):
        """conv_in (ControlNet: + control_conv_in, one merged conv) with the statistics for the first norm1"""
        B, T, H, W = g
        st = aux.gn_stats(B * T, H * W, self.cfg["block_out_channels"][0])
        return ops.conv3x3(inp64, B * T, H, W, self.conv_in_w, bias=self.conv_in_b, gn=_gn(st)), st

    def _encode(self, x, st, aux, g):
        """conv_in output (+ its GroupNorm statistics) -> (mid input, its statistics, skip list, geometry list).
        Every stage is handed the statistics table of the GroupNorm that consumes its output (None before a
        downsampler, whose conv fills the table instead) so that no norm of the encoder needs a statistics pass."""
        B, T, H, W = g
        F_ = B * T
        skips, geoms = [x], [g]
        for res, att, ds in self.down:
            for j, r in enumerate(res):
                last = j == len(res) - 1
                HW = g[2] * g[3]
                st_blk = None if (last and ds is not None) else aux.gn_stats(F_, HW, r.cout)
                if att is not None:
                    st_a = aux.gn_stats(F_, HW, r.cout)
                    x = r(x, aux, g, st_in=st, st_out=st_a)
                    x = att[j](x, aux, g, st_in=st_a, st_out=st_blk)
                else:
                    x = r(x, aux, g, st_in=st, st_out=st_blk)
                st = st_blk
                skips.append(x); geoms.append(g)
            if ds is not None:
                g = (B, T, g[2] // 2, g[3] // 2)
                st = aux.gn_stats(F_, g[2] * g[3], x.shape[1])
                x = ops.conv3x3(x, F_, g[2] * 2, g[3] * 2, ds.w, stride=2, bias=ds.b, gn=_gn(st))
                skips.append(x); geoms.append(g)
        return x, st, skips, geoms, g

    def _mid(self, x, st, aux, g, st_out=None):
        r0, a0, r1 = self.mid
        F_, HW = g[0] * g[1], g[2] * g[3]
        st_a = aux.gn_stats(F_, HW, r0.cout)
        x = r0(x, aux, g, st_in=st, st_out=st_a)
        st_r = aux.gn_stats(F_, HW, r0.cout)
        x = a0(x, aux, g, st_in=st_a, st_out=st_r)
        return r1(x, aux, g, st_in=st_r, st_out=st_out)

    def _all_blocks(self):
        res, att = [], []
        for r, a, _ in self.down:
            res += r
            att += a or []
        res += [self.mid[0], self.mid[2]]
        att.append(self.mid[1])
        return res, att


def _to_rows(t: torch.Tensor, C: int) -> torch.Tensor:
    """Reference-layout tensor [F, C, h, w] -> channels-last rows [F*h*w, C] bf16 (zero-copy when the
    tensor is already a channels-last view produced by this package)."""
    F_, Cc, h, w = t.shape
    assert Cc == C
    p = t.permute(0, 2, 3, 1)
    if t.dtype == BF16 and p.is_contiguous():
        return p.reshape(F_ * h * w, C)
    return p.to(BF16).contiguous().reshape(F_ * h * w, C)


def _from_rows(r: torch.Tensor, F_: int, h: int, w: int) -> torch.Tensor:
    """rows [F*h*w, C] -> logical [F, C, h, w] view (channels-last strides, zero-copy)."""
    return r.view(F_, h, w, r.shape[1]).permute(0, 3, 1, 2)


class ControlNetModel(_PackedModel):
    """Drop-in for ctrlv.models.ControlNetModel (controlnet.py:20-351) on sm_100a kernels."""

    is_controlnet = True

    def _pack(self):
        super()._pack()
        sd = self._sd
        k = 0
        self.zero_convs = []
        while f"controlnet_down_blocks.{k}.weight" in sd:
            self.zero_convs.append(_Lin(sd, f"controlnet_down_blocks.{k}"))
            k += 1
        self.zero_mid = _Lin(sd, "controlnet_mid_block")
        self._finish_pack(*self._all_blocks())

    @classmethod
    def from_unet(cls, unet: "UNetSpatioTemporalConditionModel", load_weights_from_unet: bool = True):
        c = unet.cfg  # controlnet.py:197-224
        over = {k: c[k] for k in c if k not in ("out_channels", "up_block_types")}
        ctrl = cls(time_context_order=unet.time_context_order, **over)
        if load_weights_from_unet:
            sd = ctrl.state_dict()
            usd = unet.state_dict()
            for k in sd:
                if k in usd:
                    sd[k] = usd[k].clone()
            ctrl.load_state_dict(sd)
        return ctrl

    def forward_rows(self, inp64, emb, ehs, g, conditioning_scale: float = 1.0, branch: Optional[int] = None):
        """inp64: [M, 64] padded channels-last input [sample(8) | control_cond(4) | 0]."""
        B, T, H, W = g
        aux = self._aux(emb, ehs, branch, n_units=B * T)
        x, st = self._conv_in(inp64, aux, g)
        x, st, skips, geoms, gm = self._encode(x, st, aux, g)
        x = self._mid(x, st, aux, gm)
        res = [ops.linear(s, z.w, bias=z.b, s_acc=float(conditioning_scale)) for s, z in zip(skips, self.zero_convs)]
        mid = ops.linear(x, self.zero_mid.w, bias=self.zero_mid.b, s_acc=float(conditioning_scale))
        return res, mid, geoms, gm

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, control_cond=None,
                conditioning_scale: float = 1.0, return_dict: bool = True, controlnet_cond=None):
        if control_cond is None:
            control_cond = controlnet_cond  # the docstring name in controlnet.py:249
        self._check_inputs(sample, timestep, encoder_hidden_states, added_time_ids)
        if control_cond is None:
            raise ValueError("`control_cond` is required (controlnet.py:289 flattens it unconditionally)")
        B, T, Cin, H, W = sample.shape
        if Cin != self.cfg["in_channels"] or control_cond.shape[2] != self.cfg["in_channels"] // 2:
            raise ValueError(f"expected sample with {self.cfg['in_channels']} and control_cond with "
                             f"{self.cfg['in_channels'] // 2} channels")
        out_dtype = sample.dtype
        inp = torch.zeros((B * T * H * W, 64), device="cuda", dtype=BF16)
        ops.nchw_to_nhwc(_as_f32_or_bf16(sample).reshape(B * T, Cin, H, W), inp, 0)
        ops.nchw_to_nhwc(_as_f32_or_bf16(control_cond).reshape(B * T, Cin // 2, H, W), inp, Cin)
        emb = self._embed(sample, timestep, added_time_ids)
        ehs = encoder_hidden_states.to(device="cuda", dtype=torch.float32).contiguous()  # [B, L, D]
        res, mid, geoms, gm = self.forward_rows(inp, emb, ehs, (B, T, H, W), conditioning_scale)
        down = [_from_rows(r, B * T, gg[2], gg[3]) for r, gg in zip(res, geoms)]
        midt = _from_rows(mid, B * T, gm[2], gm[3])
        if out_dtype != BF16:
            down = [d.to(out_dtype) for d in down]
            midt = midt.to(out_dtype)
        if not return_dict:
            return (down, midt)
        return _Output(down_block_res_samples=down, mid_block_res_sample=midt)

    __call__ = forward


def _as_f32_or_bf16(t: torch.Tensor) -> torch.Tensor:
    t = t.to("cuda")
    if t.dtype not in (torch.float32, BF16):
        t = t.float()
    return t.contiguous()


class UNetSpatioTemporalConditionModel(_PackedModel):
    """Drop-in for ctrlv.models.UNetSpatioTemporalConditionModel
    (unet_spatio_temporal_condition.py:13-171) on sm_100a kernels."""

    is_controlnet = False

    def _pack(self):
        super()._pack()
        sd, cfg = self._sd, self.cfg
        boc = tuple(cfg["block_out_channels"])
        n = len(boc)
        rheads = _tup(cfg["num_attention_heads"], n)[::-1]
        rlpb = _tup(cfg["layers_per_block"], n)[::-1]
        order = self.time_context_order
        self.up = []
        for i, t in enumerate(cfg["up_block_types"]):
            attn = t.startswith("CrossAttn")
            nl = rlpb[i] + 1
            res = [_ResBlock(sd, f"up_blocks.{i}.resnets.{j}", 1e-6) for j in range(nl)]
            att = [_Transformer(sd, f"up_blocks.{i}.attentions.{j}", rheads[i], order) for j in range(nl)] if attn else None
            us = None
            if i != n - 1:
                us = _Lin(sd, f"up_blocks.{i}.upsamplers.0.conv")
                us.w = ops.pack_upconv3x3(sd[f"up_blocks.{i}.upsamplers.0.conv.weight"])  # four 2x2 phase kernels
            self.up.append((res, att, us))
        self.norm_out = _Norm(sd, "conv_norm_out")
        oc = cfg["out_channels"]
        assert oc <= 32
        w = torch.zeros(32, 9 * boc[0], device=sd["conv_out.weight"].device, dtype=torch.float32)
        w[:oc] = _conv9(sd["conv_out.weight"].float())
        b = torch.zeros(32, device=sd["conv_out.weight"].device, dtype=torch.float32)
        b[:oc] = sd["conv_out.bias"].float()
        self.conv_out_w, self.conv_out_b = _w(w), _f(b)
        res, att = self._all_blocks()
        for r, a, _ in self.up:
            res += r
            att += a or []
        self._finish_pack(res, att)

    def forward_rows(self, inp64, emb, ehs, g, down_res=None, mid_res=None, out_f32=None, join=None,
                     branch: Optional[int] = None):
        """inp64 [M, 64] -> noise prediction rows [M, out_channels] fp32."""
        B, T, H, W = g
        aux = self._aux(emb, ehs, branch, n_units=B * T)
        F_ = B * T
        x, st = self._conv_in(inp64, aux, g)
        x, st, skips, geoms, gm = self._encode(x, st, aux, g)
        if join is not None:
            join()  # the residuals come from another stream (DenoiseStep two-stream mode)
        # Decoder ResBlock k (in execution order) normalises cat([x (cx[k] channels), skip k]): the stage before
        # it accumulates the x half of table k at channel offset 0, the residual add (skip + r) the skip half at
        # offset cx[k].  Without residuals a raw skip already fed an encoder norm (a second consumer, another group
        # width): its half comes from one statistics-only pass of the same reduction, so that a zero residual and
        # no residual give bit-identical results.
        cx, c = [], skips[-1].shape[1]
        for res, _, _ in self.up:
            for r in res:
                cx.append(c)
                c = r.cout
        nsk = len(skips)
        st_cat = [None] * (nsk + 1)  # k-th pop <-> skips[nsk - 1 - k]; [nsk]: no further norm1
        for i, (s_, gg) in enumerate(zip(skips, geoms)):
            st_cat[nsk - 1 - i] = aux.gn_stats(F_, gg[2] * gg[3], cx[nsk - 1 - i] + s_.shape[1])
        if down_res is not None:  # unet_spatio_temporal_condition.py:119-127
            skips = [ops.axpby(s_, r_, gn=_gn(st_cat[nsk - 1 - i], cx[nsk - 1 - i]))
                     for i, (s_, r_) in enumerate(zip(skips, down_res))]
        else:
            for i, s_ in enumerate(skips):
                if st_cat[nsk - 1 - i] is not None:
                    ops.gn_stats_of(s_, (st_cat[nsk - 1 - i], cx[nsk - 1 - i]))
        x = self._mid(x, st, aux, gm)
        if mid_res is not None:  # :136-137
            x = ops.axpby(x, mid_res, gn=_gn(st_cat[0], 0))
        elif st_cat[0] is not None:
            ops.gn_stats_of(x, (st_cat[0], 0))  # (same reduction as the residual add: see above)
        g = gm
        k = 0
        st = None
        for i, (res, att, us) in enumerate(self.up):
            HW = g[2] * g[3]
            for j, r in enumerate(res):
                skip = skips.pop()
                if j < len(res) - 1:
                    st_stage = st_cat[k + 1]        # the next ResBlock of this block
                elif us is not None:
                    st_stage = None                 # the upsampler's convs fill table k + 1
                else:
                    st_stage = aux.gn_stats(F_, HW, r.cout)  # conv_norm_out
                if att is not None:
                    st_a = aux.gn_stats(F_, HW, r.cout)
                    x = r(x, aux, g, x1=skip, st_in=st_cat[k], st_out=st_a)
                    x = att[j](x, aux, g, st_in=st_a, st_out=st_stage)
                else:
                    x = r(x, aux, g, x1=skip, st_in=st_cat[k], st_out=st_stage)
                st = st_stage
                k += 1
            if us is not None:
                # Upsample2D: nearest 2x + 3x3 conv, fused as four 2x2 phase convs of the low-res frame
                x = ops.upsample2x_conv3x3(x, F_, g[2], g[3], us.w, bias=us.b, gn=_gn(st_cat[k]))
                g = (B, T, g[2] * 2, g[3] * 2)
        a = ops.groupnorm(x, B * T, g[2] * g[3], self.norm_out.g, self.norm_out.b, 1e-5, True, stats=st)
        oc = self.cfg["out_channels"]
        if out_f32 is None:
            out_f32 = torch.empty((x.shape[0], oc), device="cuda", dtype=torch.float32)
        ops.conv3x3(a, B * T, g[2], g[3], self.conv_out_w, bias=self.conv_out_b, out_f32=out_f32, n_store=oc)
        return out_f32

    @torch.no_grad()
    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids,
                down_block_additional_residuals=None, mid_block_additional_residuals=None,
                return_dict: bool = True):
        self._check_inputs(sample, timestep, encoder_hidden_states, added_time_ids)
        B, T, Cin, H, W = sample.shape
        if Cin != self.cfg["in_channels"]:
            raise ValueError(f"expected sample with {self.cfg['in_channels']} channels, got {Cin}")
        is_controlnet = mid_block_additional_residuals is not None and down_block_additional_residuals is not None
        out_dtype = sample.dtype
        inp = torch.zeros((B * T * H * W, 64), device="cuda", dtype=BF16)
        ops.nchw_to_nhwc(_as_f32_or_bf16(sample).reshape(B * T, Cin, H, W), inp, 0)
        emb = self._embed(sample, timestep, added_time_ids)
        ehs = encoder_hidden_states.to(device="cuda", dtype=torch.float32).contiguous()  # [B, L, D]
        down = mid = None
        if is_controlnet:
            down = [_to_rows(r, r.shape[1]) for r in down_block_additional_residuals]
            mid = _to_rows(mid_block_additional_residuals, mid_block_additional_residuals.shape[1])
        rows = self.forward_rows(inp, emb, ehs, (B, T, H, W), down, mid)
        oc = self.cfg["out_channels"]
        out = ops.nhwc_to_nchw(rows, B * T, oc, H, W, dtype=torch.float32).reshape(B, T, oc, H, W)
        if out_dtype != torch.float32:
            out = out.to(out_dtype)
        if not return_dict:
            return (out,)
        return _Output(sample=out)

    __call__ = forward
