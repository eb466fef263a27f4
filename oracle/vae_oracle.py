"""ORACLE (test infrastructure — never imported by the product path).

CPU/fp32 restatement of diffusers==0.27.2 `AutoencoderKLTemporalDecoder`
(models/autoencoders/autoencoder_kl_temporal_decoder.py, models/autoencoders/vae.py `Encoder`,
models/unets/unet_3d_blocks.py `MidBlockTemporalDecoder` / `UpBlockTemporalDecoder`,
models/unets/unet_2d_blocks.py `DownEncoderBlock2D` / `UNetMidBlock2D`, models/attention_processor.py
`Attention` with `group_norm` + `residual_connection`), which is the `vae` of the reference pipelines:
  encode  /root/reference/src/ctrlv/pipelines/pipeline_video_control.py:84 (`latent_dist.mode()`), :235
  decode  /root/reference/src/ctrlv/pipelines/pipeline_video_control.py:346-347 (`decode_latents`,
          diffusers pipeline_stable_video_diffusion.py) and pipeline_video_diffusion.py:50-54,296.
diffusers is absent from this image, so this follows SURVEY.md A.11 and the published module
structure, with the diffusers state-dict key names.  PARITY UNPINNED: no reference fixture exists
for the VAE; pinned only by structure (key names, shapes) and algebraic checks in tests/.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .svd_oracle import ResnetBlock2D, SpatioTemporalResBlock, Upsample2D

VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(128, 256, 512, 512),
                  layers_per_block=2, scaling_factor=0.18215, force_upcast=True)
TINY_VAE_CONFIG = dict(in_channels=3, out_channels=3, latent_channels=4, block_out_channels=(64, 128),
                       layers_per_block=1, scaling_factor=0.18215, force_upcast=True)


class VaeAttention(nn.Module):
    """`Attention(C, heads=C//dim_head, dim_head, norm_num_groups=32, eps=1e-6, bias=True,
    residual_connection=True, _from_deprecated_attn_block=True)` on a [N, C, H, W] map."""

    def __init__(self, channels: int, dim_head: int):
        super().__init__()
        self.heads = channels // dim_head
        self.group_norm = nn.GroupNorm(32, channels, eps=1e-6, affine=True)
        self.to_q = nn.Linear(channels, channels, bias=True)
        self.to_k = nn.Linear(channels, channels, bias=True)
        self.to_v = nn.Linear(channels, channels, bias=True)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])

    def forward(self, x):
        residual = x
        n, c, h, w = x.shape
        hs = x.view(n, c, h * w).transpose(1, 2)
        hs = self.group_norm(hs.transpose(1, 2)).transpose(1, 2)
        q, k, v = self.to_q(hs), self.to_k(hs), self.to_v(hs)
        d = c // self.heads
        q, k, v = (t.view(n, -1, self.heads, d).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(n, -1, c)
        o = self.to_out[0](o)
        o = o.transpose(-1, -2).reshape(n, c, h, w)
        return o + residual  # rescale_output_factor = 1


class EncoderDownsample(nn.Module):
    """Downsample2D(use_conv=True, padding=0): pad (0,1,0,1) then 3x3 stride-2 conv without padding."""

    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, cin, cout, num_layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, 1e-6)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([EncoderDownsample(cout)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels: int, dim_head: int):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, None, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([VaeAttention(channels, dim_head)])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class Encoder(nn.Module):
    """vae.py `Encoder(double_z=True)` with DownEncoderBlock2D blocks."""

    def __init__(self, in_channels, latent_channels, block_out_channels, layers_per_block):
        super().__init__()
        boc = block_out_channels
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, co in enumerate(boc):
            self.down_blocks.append(DownEncoderBlock2D(c, co, layers_per_block, i != len(boc) - 1))
            c = co
        self.mid_block = UNetMidBlock2D(boc[-1], boc[-1])
        self.conv_norm_out = nn.GroupNorm(32, boc[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[-1], 2 * latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def _st_block(cin, cout):
    return SpatioTemporalResBlock(cin, cout, None, eps=1e-6, temporal_eps=1e-5, merge_factor=0.0,
                                  merge_strategy="learned", switch_spatial_to_temporal_mix=True)


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, channels: int, dim_head: int, num_layers: int):
        super().__init__()
        self.resnets = nn.ModuleList([_st_block(channels, channels) for _ in range(num_layers)])
        self.attentions = nn.ModuleList([VaeAttention(channels, dim_head)])

    def forward(self, x, image_only_indicator):
        x = self.resnets[0](x, None, image_only_indicator)
        for resnet, attn in zip(self.resnets[1:], self.attentions):
            x = attn(x)
            x = resnet(x, None, image_only_indicator)
        return x


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, cin, cout, num_layers, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_st_block(cin if i == 0 else cout, cout) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_upsample else None

    def forward(self, x, image_only_indicator):
        for r in self.resnets:
            x = r(x, None, image_only_indicator)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        boc = block_out_channels
        self.conv_in = nn.Conv2d(in_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(boc[-1], boc[-1], layers_per_block)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(boc))
        c = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(UpBlockTemporalDecoder(c, co, layers_per_block + 1, i != len(boc) - 1))
            c = co
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-6)
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, image_only_indicator, num_frames: int = 1):
        x = self.conv_in(z)
        x = self.mid_block(x, image_only_indicator)
        for b in self.up_blocks:
            x = b(x, image_only_indicator)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        b = bf // num_frames
        x = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class AutoencoderKLTemporalDecoder(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(VAE_CONFIG)
        cfg.update(overrides)
        self.cfg = cfg
        self.encoder = Encoder(cfg["in_channels"], cfg["latent_channels"], cfg["block_out_channels"],
                               cfg["layers_per_block"])
        self.decoder = TemporalDecoder(cfg["latent_channels"], cfg["out_channels"], cfg["block_out_channels"],
                                       cfg["layers_per_block"])
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)

    def encode_moments(self, x):
        return self.quant_conv(self.encoder(x))

    def encode_mode(self, x):
        """`vae.encode(x).latent_dist.mode()`: the mean half of the moments."""
        return self.encode_moments(x).chunk(2, dim=1)[0]

    def decode(self, z, num_frames: int):
        b = z.shape[0] // num_frames
        ioi = torch.zeros(b, num_frames, dtype=z.dtype, device=z.device)
        return self.decoder(z, ioi, num_frames=num_frames)


def decode_latents(vae: AutoencoderKLTemporalDecoder, latents, num_frames: int, decode_chunk_size: int = 14):
    """StableVideoDiffusionPipeline.decode_latents: [B, T, 4, h, w] -> [B, 3, T, H, W] float."""
    latents = latents.flatten(0, 1)
    latents = 1 / vae.cfg["scaling_factor"] * latents
    frames = []
    for i in range(0, latents.shape[0], decode_chunk_size):
        chunk = latents[i:i + decode_chunk_size]
        frames.append(vae.decode(chunk, num_frames=chunk.shape[0]))
    frames = torch.cat(frames, dim=0)
    frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
    return frames.float()
