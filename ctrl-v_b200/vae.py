"""Temporal VAE of the Box2Video pipelines on the sm_100a kernels.

Drop-in for diffusers' `AutoencoderKLTemporalDecoder` as the reference pipelines use it:
  * `vae.encode(x).latent_dist.mode()`   pipeline_video_control.py:84 (bbox frames), :235 (image)
  * `vae.decode(z, num_frames=n).sample` via `decode_latents`, pipeline_video_control.py:346-347,
                                          pipeline_video_diffusion.py:50-54,296
  * `vae.config.scaling_factor`, `vae.config.force_upcast`, `vae.dtype`

Same building blocks as the denoiser: channels-last bf16 rows, tcgen05 implicit-GEMM convolutions
(3x3, 1x1 shortcut fused into conv2's K loop, (3,1,1) temporal), two-pass GroupNorm(+SiLU).  New for
the VAE: the encoder's asymmetric-pad stride-2 conv (four parity sub-lattices), the 512-wide
single-head mid-block attention as two GEMMs around a row softmax, and `time_conv_out` fused with
the NCHW conversion.  `quant_conv` (1x1) is folded into the encoder's `conv_out` at pack time, and
`.mode()` keeps only the mean rows.  No PyTorch math on the path.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .models import BF16, _Lin, _Norm, _conv9, _conv_t3, _f, _w

VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, scaling_factor=0.18215, force_upcast=True)


# ------------------------------------------------------------------------------------------------
# parameter spec (diffusers key names) and random init
# ------------------------------------------------------------------------------------------------
def _spec_conv(spec, name, cin, cout, k=3):
    spec[name + ".weight"] = (cout, cin, k, k); spec[name + ".bias"] = (cout,)


def _spec_norm(spec, name, c):
    spec[name + ".weight"] = (c,); spec[name + ".bias"] = (c,)


def _spec_res2d(spec, pfx, cin, cout):
    _spec_norm(spec, pfx + ".norm1", cin); _spec_conv(spec, pfx + ".conv1", cin, cout)
    _spec_norm(spec, pfx + ".norm2", cout); _spec_conv(spec, pfx + ".conv2", cout, cout)
    if cin != cout:
        _spec_conv(spec, pfx + ".conv_shortcut", cin, cout, 1)


def _spec_st(spec, pfx, cin, cout):
    _spec_res2d(spec, pfx + ".spatial_res_block", cin, cout)
    t = pfx + ".temporal_res_block"
    for n in ("1", "2"):
        _spec_norm(spec, t + ".norm" + n, cout)
        spec[t + f".conv{n}.weight"] = (cout, cout, 3, 1, 1); spec[t + f".conv{n}.bias"] = (cout,)
    spec[pfx + ".time_mixer.mix_factor"] = (1,)


def _spec_attn(spec, pfx, c):
    _spec_norm(spec, pfx + ".group_norm", c)
    for n in ("to_q", "to_k", "to_v", "to_out.0"):
        spec[pfx + f".{n}.weight"] = (c, c); spec[pfx + f".{n}.bias"] = (c,)


def param_spec(cfg: dict) -> "OrderedDict[str, tuple]":
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    boc, lpb, lc = tuple(cfg["block_out_channels"]), cfg["layers_per_block"], cfg["latent_channels"]
    _spec_conv(spec, "encoder.conv_in", cfg["in_channels"], boc[0])
    c = boc[0]
    for i, co in enumerate(boc):
        for j in range(lpb):
            _spec_res2d(spec, f"encoder.down_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i != len(boc) - 1:
            _spec_conv(spec, f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co)
        c = co
    _spec_res2d(spec, "encoder.mid_block.resnets.0", c, c)
    _spec_attn(spec, "encoder.mid_block.attentions.0", c)
    _spec_res2d(spec, "encoder.mid_block.resnets.1", c, c)
    _spec_norm(spec, "encoder.conv_norm_out", c); _spec_conv(spec, "encoder.conv_out", c, 2 * lc)
    _spec_conv(spec, "decoder.conv_in", lc, boc[-1])
    for j in range(lpb):
        _spec_st(spec, f"decoder.mid_block.resnets.{j}", boc[-1], boc[-1])
    _spec_attn(spec, "decoder.mid_block.attentions.0", boc[-1])
    rev = list(reversed(boc))
    c = rev[0]
    for i, co in enumerate(rev):
        for j in range(lpb + 1):
            _spec_st(spec, f"decoder.up_blocks.{i}.resnets.{j}", c if j == 0 else co, co)
        if i != len(boc) - 1:
            _spec_conv(spec, f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co)
        c = co
    _spec_norm(spec, "decoder.conv_norm_out", boc[0]); _spec_conv(spec, "decoder.conv_out", boc[0], cfg["out_channels"])
    oc = cfg["out_channels"]
    spec["decoder.time_conv_out.weight"] = (oc, oc, 3, 1, 1); spec["decoder.time_conv_out.bias"] = (oc,)
    _spec_conv(spec, "quant_conv", 2 * lc, 2 * lc, 1)
    return spec


def random_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator("cpu").manual_seed(seed)
    sd = OrderedDict()
    for k, shape in param_spec(cfg).items():
        if k.endswith("mix_factor"):
            sd[k] = torch.zeros(shape)
        elif ".norm" in k or "group_norm" in k or "conv_norm_out" in k:
            sd[k] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if k.endswith("weight") else 0.1 * torch.randn(shape, generator=g)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            sd[k] = torch.randn(shape, generator=g) / fan_in ** 0.5
    return sd


# ------------------------------------------------------------------------------------------------
# packed blocks
# ------------------------------------------------------------------------------------------------
def _pad_cin(w: torch.Tensor, cpad: int) -> torch.Tensor:  # [Cout, Cin, 3, 3] -> Cin zero-padded
    out = torch.zeros((w.shape[0], cpad, 3, 3), dtype=torch.float32, device=w.device)
    out[:, :w.shape[1]] = w.float()
    return out


class _Res2D:
    """ResnetBlock2D without a time embedding (eps 1e-6): GN-SiLU-conv, GN-SiLU-conv (+1x1 shortcut
    riding in conv2's K loop) + x."""

    def __init__(self, sd, pfx, eps=1e-6):
        self.eps = eps
        self.norm1, self.norm2 = _Norm(sd, pfx + ".norm1"), _Norm(sd, pfx + ".norm2")
        w1 = sd[pfx + ".conv1.weight"]
        self.cin, self.cout = w1.shape[1], w1.shape[0]
        self.conv1_w, self.conv1_b = _w(_conv9(w1.float())), _f(sd[pfx + ".conv1.bias"])
        w2 = _conv9(sd[pfx + ".conv2.weight"].float())
        b2 = sd[pfx + ".conv2.bias"].float()
        self.has_shortcut = (pfx + ".conv_shortcut.weight") in sd
        if self.has_shortcut:
            ws = sd[pfx + ".conv_shortcut.weight"].float().reshape(self.cout, self.cin)
            w2 = torch.cat([w2, ws.to(w2.device)], dim=1)
            b2 = b2 + sd[pfx + ".conv_shortcut.bias"].float().to(b2.device)
        self.conv2_w, self.conv2_b = _w(w2), _f(b2)

    def __call__(self, x, F_, H, W):
        a1 = ops.groupnorm(x, F_, H * W, self.norm1.g, self.norm1.b, self.eps, True)
        h = ops.conv3x3(a1, F_, H, W, self.conv1_w, bias=self.conv1_b)
        a2 = ops.groupnorm(h, F_, H * W, self.norm2.g, self.norm2.b, self.eps, True)
        if self.has_shortcut:
            return ops.conv3x3(a2, F_, H, W, self.conv2_w, sc0=x, bias=self.conv2_b)
        return ops.conv3x3(a2, F_, H, W, self.conv2_w, bias=self.conv2_b, res1=x)


class _STRes:
    """SpatioTemporalResBlock(temb=None, eps 1e-6, temporal_eps 1e-5, merge "learned",
    switch_spatial_to_temporal_mix=True): alpha = 1 - sigmoid(mix);
    out = alpha*xs + (1-alpha)*(xs + h_t) = xs + sigmoid(mix)*h_t."""

    def __init__(self, sd, pfx):
        self.spatial = _Res2D(sd, pfx + ".spatial_res_block", 1e-6)
        t = pfx + ".temporal_res_block"
        self.teps = 1e-5
        self.tnorm1, self.tnorm2 = _Norm(sd, t + ".norm1"), _Norm(sd, t + ".norm2")
        self.tconv1_w, self.tconv1_b = _w(_conv_t3(sd[t + ".conv1.weight"].float())), _f(sd[t + ".conv1.bias"])
        self.tconv2_w, self.tconv2_b = _w(_conv_t3(sd[t + ".conv2.weight"].float())), _f(sd[t + ".conv2.bias"])
        self.beta = float(torch.sigmoid(sd[pfx + ".time_mixer.mix_factor"].float()).item())

    def __call__(self, x, B, T, H, W):
        HW = H * W
        xs = self.spatial(x, B * T, H, W)
        a3 = ops.groupnorm(xs, B, T * HW, self.tnorm1.g, self.tnorm1.b, self.teps, True)
        h2 = ops.conv_t3(a3, B, T, HW, self.tconv1_w, bias=self.tconv1_b)
        a4 = ops.groupnorm(h2, B, T * HW, self.tnorm2.g, self.tnorm2.b, self.teps, True)
        return ops.conv_t3(a4, B, T, HW, self.tconv2_w, bias=self.tconv2_b, s_acc=self.beta, res1=xs, s_res1=1.0)


class _Attn:
    """Single-head attention of the VAE mid blocks (dim_head == channels): GroupNorm, q/k/v, softmax
    (q k^T / sqrt(C)) v, to_out, + residual.  Per frame: scores GEMM (fp32 out) -> row softmax ->
    P V GEMM against V^T, which is produced directly by a GEMM with swapped operands; the value
    bias is added after P V (rows of P sum to one)."""

    def __init__(self, sd, pfx):
        self.norm = _Norm(sd, pfx + ".group_norm")
        self.q, self.k = _Lin(sd, pfx + ".to_q"), _Lin(sd, pfx + ".to_k")
        self.v, self.o = _Lin(sd, pfx + ".to_v"), _Lin(sd, pfx + ".to_out.0")
        self.C = self.q.w.shape[0]

    def __call__(self, x, F_, S):
        Cc = self.C
        a = ops.groupnorm(x, F_, S, self.norm.g, self.norm.b, 1e-6, False)
        q = ops.linear(a, self.q.w, bias=self.q.b)
        k = ops.linear(a, self.k.w, bias=self.k.b)
        Sp = (S + 63) // 64 * 64  # GEMM N and K extents are multiples of 64: zero-pad ragged S
        o = torch.empty((F_ * S, Cc), dtype=BF16, device="cuda")
        scores = torch.empty((S, Sp), dtype=torch.float32, device="cuda")
        probs = torch.zeros((S, Sp), dtype=BF16, device="cuda")
        vt = torch.empty((Cc, Sp), dtype=BF16, device="cuda")
        ap = kp = None
        if Sp != S:
            ap = torch.zeros((Sp, Cc), dtype=BF16, device="cuda")
            kp = torch.zeros((Sp, Cc), dtype=BF16, device="cuda")
        scale = float(Cc) ** -0.5
        for f in range(F_):
            rows = slice(f * S, (f + 1) * S)
            af, kf = a[rows], k[rows]
            if Sp != S:
                ap[:S].copy_(af); kp[:S].copy_(kf)
                af, kf = ap, kp
            ops.linear(self.v.w, af, out=vt)                        # V^T (without bias) [C, Sp]
            ops.linear(q[rows], kf, out_f32=scores)                 # q k^T [S, Sp]
            ops.softmax_rows(scores[:, :S], scale, out=probs[:, :S])
            ops.linear(probs, vt, bias=self.v.b, out=o[rows])      # P V + b_v
        return ops.linear(o, self.o.w, bias=self.o.b, res1=x)


class _DiagonalGaussian(SimpleNamespace):
    def mode(self):
        return self.mean


class AutoencoderKLTemporalDecoder(torch.nn.Module):
    """Same constructor convention as the denoiser drop-ins: `state_dict` with diffusers key names
    (random init when omitted), config overrides as keyword arguments."""

    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0, **overrides):
        super().__init__()
        cfg = dict(VAE_CONFIG)
        cfg.update(overrides)
        cfg["block_out_channels"] = tuple(cfg["block_out_channels"])
        for c in cfg["block_out_channels"]:
            if c % 64 != 0:
                raise ValueError(f"block_out_channels must be multiples of 64 for the sm_100a kernels, got {c}")
        if cfg["out_channels"] > 4 or cfg["in_channels"] > 64 or cfg["latent_channels"] > 16:
            raise ValueError("unsupported VAE channel configuration")
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.dtype = BF16
        self._sd: Dict[str, torch.Tensor] = {}
        self.load_state_dict(state_dict if state_dict is not None else random_state_dict(cfg, seed))

    # ---- state dict / checkpoints ---------------------------------------------------------------
    def state_dict(self, *a, **k):
        return OrderedDict(self._sd)

    def load_state_dict(self, sd, strict: bool = True):
        spec = param_spec(self.cfg)
        missing = [k for k in spec if k not in sd]
        unexpected = [k for k in sd if k not in spec]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:5]} unexpected {unexpected[:5]}")
        for k, shape in spec.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(sd[k].shape)} vs {shape}")
                self._sd[k] = sd[k].detach().to("cuda")
        self._pack()
        return SimpleNamespace(missing_keys=missing, unexpected_keys=unexpected)

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, variant: Optional[str] = None, **kwargs):
        from . import checkpoint
        kwargs.pop("torch_dtype", None)
        config, sd = checkpoint.load_diffusers_dir(path, subfolder, variant)
        over = {k: (tuple(v) if isinstance(v, list) else v) for k, v in config.items() if k in VAE_CONFIG}
        over.update(kwargs)
        return cls(state_dict=sd, **over)

    def save_pretrained(self, save_directory: str, subfolder: Optional[str] = None, variant: Optional[str] = None,
                        safe_serialization: bool = True):
        from . import checkpoint
        checkpoint.save_diffusers_dir(save_directory, self.cfg, self._sd, "AutoencoderKLTemporalDecoder",
                                      subfolder, variant, safe_serialization)

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    # ---- packing ----------------------------------------------------------------------------------
    def _pack(self):
        sd, cfg = self._sd, self.cfg
        boc, lpb, lc = cfg["block_out_channels"], cfg["layers_per_block"], cfg["latent_channels"]
        # encoder
        self.e_conv_in_w = _w(_conv9(_pad_cin(sd["encoder.conv_in.weight"], 64)))
        self.e_conv_in_b = _f(sd["encoder.conv_in.bias"])
        self.e_down = []
        for i in range(len(boc)):
            res = [_Res2D(sd, f"encoder.down_blocks.{i}.resnets.{j}") for j in range(lpb)]
            ds = None
            if i != len(boc) - 1:
                n = f"encoder.down_blocks.{i}.downsamplers.0.conv"
                ds = (_w(_conv9(sd[n + ".weight"].float())), _f(sd[n + ".bias"]))
            self.e_down.append((res, ds))
        self.e_mid = (_Res2D(sd, "encoder.mid_block.resnets.0"), _Attn(sd, "encoder.mid_block.attentions.0"),
                      _Res2D(sd, "encoder.mid_block.resnets.1"))
        self.e_norm_out = _Norm(sd, "encoder.conv_norm_out")
        # conv_out followed by quant_conv (1x1): fold; .mode() keeps the mean rows only
        wq = sd["quant_conv.weight"].float().reshape(2 * lc, 2 * lc)
        wc = sd["encoder.conv_out.weight"].float()
        wf = torch.einsum("om,mcyx->ocyx", wq, wc)
        bf_ = wq @ sd["encoder.conv_out.bias"].float() + sd["quant_conv.bias"].float()
        w = torch.zeros((32, 9 * boc[-1]), dtype=torch.float32, device=wf.device)
        b = torch.zeros((32,), dtype=torch.float32, device=wf.device)
        w[:2 * lc] = _conv9(wf); b[:2 * lc] = bf_
        self.e_conv_out_w, self.e_conv_out_b = _w(w), _f(b)
        # decoder
        self._d_conv_in_scaled = {1.0: _w(_conv9(_pad_cin(sd["decoder.conv_in.weight"], 64)))}
        self.d_conv_in_b = _f(sd["decoder.conv_in.bias"])
        self.d_mid_res = [_STRes(sd, f"decoder.mid_block.resnets.{j}") for j in range(lpb)]
        self.d_mid_attn = _Attn(sd, "decoder.mid_block.attentions.0")
        self.d_up = []
        for i in range(len(boc)):
            res = [_STRes(sd, f"decoder.up_blocks.{i}.resnets.{j}") for j in range(lpb + 1)]
            us = None
            if i != len(boc) - 1:
                n = f"decoder.up_blocks.{i}.upsamplers.0.conv"
                us = (ops.pack_upconv3x3(sd[n + ".weight"]), _f(sd[n + ".bias"]))
            self.d_up.append((res, us))
        self.d_norm_out = _Norm(sd, "decoder.conv_norm_out")
        oc = cfg["out_channels"]
        w = torch.zeros((32, 9 * boc[0]), dtype=torch.float32, device=wf.device)
        b = torch.zeros((32,), dtype=torch.float32, device=wf.device)
        w[:oc] = _conv9(sd["decoder.conv_out.weight"].float()); b[:oc] = sd["decoder.conv_out.bias"].float()
        self.d_conv_out_w, self.d_conv_out_b = _w(w), _f(b)
        self.d_tconv_w = _f(sd["decoder.time_conv_out.weight"].float().reshape(oc, oc, 3))
        self.d_tconv_b = _f(sd["decoder.time_conv_out.bias"])

    # ---- forward ------------------------------------------------------------------------------------
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [N, 3, H, W] in [-1, 1] -> object with `.latent_dist.mode()` = mean [N, 4, H/f, W/f] fp32
        (f = 2^(levels-1); sampling from the posterior is not needed by the reference pipelines)."""
        if x.ndim != 4 or x.shape[1] != self.cfg["in_channels"]:
            raise ValueError(f"encode expects [N, {self.cfg['in_channels']}, H, W], got {tuple(x.shape)}")
        F_, _, H, W = x.shape
        f = 2 ** (len(self.cfg["block_out_channels"]) - 1)
        if H % f or W % f:
            raise ValueError(f"H, W must be multiples of {f}")
        src = x.to("cuda")
        src = src.contiguous() if src.dtype in (torch.float32, BF16) else src.float().contiguous()
        inp = torch.zeros((F_ * H * W, 64), dtype=BF16, device="cuda")
        ops.nchw_to_nhwc(src, inp)
        h = ops.conv3x3(inp, F_, H, W, self.e_conv_in_w, bias=self.e_conv_in_b)
        for res, ds in self.e_down:
            for r in res:
                h = r(h, F_, H, W)
            if ds is not None:
                h = ops.conv3x3_s2_pad01(h, F_, H, W, ds[0], bias=ds[1])
                H, W = H // 2, W // 2
        r0, attn, r1 = self.e_mid
        h = r1(attn(r0(h, F_, H, W), F_, H * W), F_, H, W)
        a = ops.groupnorm(h, F_, H * W, self.e_norm_out.g, self.e_norm_out.b, 1e-6, True)
        lc = self.cfg["latent_channels"]
        mom = torch.empty((F_ * H * W, 2 * lc), dtype=torch.float32, device="cuda")
        ops.conv3x3(a, F_, H, W, self.e_conv_out_w, bias=self.e_conv_out_b, out_f32=mom, n_store=2 * lc)
        mean = ops.nhwc_to_nchw(mom[:, :lc], F_, lc, H, W)
        logvar = ops.nhwc_to_nchw(mom[:, lc:], F_, lc, H, W)
        dist = _DiagonalGaussian(mean=mean, logvar=logvar)
        if not return_dict:
            return (dist,)
        return SimpleNamespace(latent_dist=dist)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, num_frames: int = 1, return_dict: bool = True, latent_scale: float = 1.0):
        """z [B*T, 4, h, w] -> `.sample` [B*T, 3, 8h, 8w] fp32 (temporal layers mix the T frames of a
        clip).  `latent_scale` multiplies z (the pipeline's 1/scaling_factor), folded into conv_in."""
        if z.ndim != 4 or z.shape[1] != self.cfg["latent_channels"]:
            raise ValueError(f"decode expects [N, {self.cfg['latent_channels']}, h, w], got {tuple(z.shape)}")
        F_, _, H, W = z.shape
        if F_ % num_frames:
            raise ValueError(f"batch {F_} is not a multiple of num_frames {num_frames}")
        B, T = F_ // num_frames, num_frames
        src = z.to("cuda")
        src = src.contiguous() if src.dtype in (torch.float32, BF16) else src.float().contiguous()
        inp = torch.zeros((F_ * H * W, 64), dtype=BF16, device="cuda")
        ops.nchw_to_nhwc(src, inp)
        w_in = self._d_conv_in_scaled.get(latent_scale)
        if w_in is None:
            w_in = _w(_conv9(_pad_cin(self._sd["decoder.conv_in.weight"], 64)) * latent_scale)
            self._d_conv_in_scaled[latent_scale] = w_in
        h = ops.conv3x3(inp, F_, H, W, w_in, bias=self.d_conv_in_b)
        h = self.d_mid_res[0](h, B, T, H, W)
        for r in self.d_mid_res[1:]:
            h = self.d_mid_attn(h, F_, H * W)
            h = r(h, B, T, H, W)
        for res, us in self.d_up:
            for r in res:
                h = r(h, B, T, H, W)
            if us is not None:
                h = ops.upsample2x_conv3x3(h, F_, H, W, us[0], bias=us[1])
                H, W = 2 * H, 2 * W
        a = ops.groupnorm(h, F_, H * W, self.d_norm_out.g, self.d_norm_out.b, 1e-6, True)
        oc = self.cfg["out_channels"]
        img = torch.empty((F_ * H * W, 4), dtype=torch.float32, device="cuda")
        ops.conv3x3(a, F_, H, W, self.d_conv_out_w, bias=self.d_conv_out_b, out_f32=img, n_store=oc)
        out = ops.time_conv_out(img, B, T, H, W, oc, self.d_tconv_w, self.d_tconv_b)
        if not return_dict:
            return (out,)
        return SimpleNamespace(sample=out)


def decode_latents(vae: AutoencoderKLTemporalDecoder, latents: torch.Tensor, num_frames: int,
                   decode_chunk_size: int = 14) -> torch.Tensor:
    """`StableVideoDiffusionPipeline.decode_latents` (called at pipeline_video_control.py:346):
    [B, T, 4, h, w] -> [B, 3, T, H, W] fp32, decoding `decode_chunk_size` frames at a time."""
    lat = latents.flatten(0, 1).to("cuda", torch.float32)
    frames = []
    for i in range(0, lat.shape[0], decode_chunk_size):
        chunk = lat[i:i + decode_chunk_size].contiguous()
        frames.append(vae.decode(chunk, num_frames=chunk.shape[0],
                                 latent_scale=1.0 / vae.config.scaling_factor).sample)
    frames = torch.cat(frames, dim=0)
    return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
