"""ORACLE (test infrastructure — never imported by the product path).

Pure-PyTorch restatement of the arithmetic behind Ctrl-V's Box2Video denoise step.  The reference
delegates all of it to the un-vendored dependency ``diffusers==0.27.2`` (requirements.txt:3), which
is absent from /root/reference and from this image, so the modules below restate that library's
published algorithm module-for-module, with the SAME state-dict key names, and
the reference's own two forwards on top:

  * ``ControlNetModel``                   <- src/ctrlv/models/controlnet.py:53-351
  * ``UNetSpatioTemporalConditionModel``  <- src/ctrlv/models/unet_spatio_temporal_condition.py:31-171
                                             (constructor: diffusers UNetSpatioTemporalConditionModel)
  * ``denoise_step`` / ``sample_loop``    <- src/ctrlv/pipelines/pipeline_video_control.py:298-343

PARITY — two layers with different status:
  * the reference's IN-REPO code (the two forwards, the ControlNet constructor / from_unet, the
    pipeline loops) is PINNED: tests/test_reference_pin.py loads those files from /root/reference as
    they lie, runs them on a stand-in `diffusers` whose blocks are the classes below
    (tests/golden/ref_shim.py) and requires bit-exact equality with this module; the vectors the
    reference code produced are committed (tests/golden/ref_forward.pt) and checked everywhere;
  * the arithmetic INSIDE the diffusers blocks (below) is UNPINNED: the reference ships no tests,
    golden vectors or fixtures and diffusers cannot be imported here, so it is
    held only by self-made checks (tests/test_oracle.py): exact parameter counts 1,524,623,082
    (UNet) and 680,946,897 (ControlNet), the diffusers key set, scheduler known answers, algebraic
    identities.

diffusers source files restated (0.27.2): models/unets/unet_spatio_temporal_condition.py,
models/unets/unet_3d_blocks.py, models/resnet.py, models/transformers/transformer_temporal.py,
models/attention.py, models/attention_processor.py, models/embeddings.py,
models/downsampling.py, models/upsampling.py.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# embeddings (diffusers models/embeddings.py)
# --------------------------------------------------------------------------------------------
class Timesteps(nn.Module):
    def __init__(self, num_channels: int, flip_sin_to_cos: bool = True, downscale_freq_shift: float = 0.0):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps: torch.Tensor) -> torch.Tensor:
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = torch.exp(exponent)
        emb = timesteps[:, None].float() * emb[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, out_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim if out_dim is not None else time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------------------
# resnets (diffusers models/resnet.py)
# --------------------------------------------------------------------------------------------
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float, groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h  # output_scale_factor = 1


class TemporalResnetBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv3d(in_channels, out_channels, (3, 1, 1), stride=1, padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv3d(out_channels, out_channels, (3, 1, 1), stride=1, padding=(1, 0, 0))
        self.conv_shortcut = (nn.Conv3d(in_channels, out_channels, 1) if in_channels != out_channels else None)

    def forward(self, x, temb):  # x [B, C, T, H, W], temb [B, T, temb_channels]
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            t = self.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
            h = h + t
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class AlphaBlender(nn.Module):
    def __init__(self, alpha: float, merge_strategy: str = "learned_with_images",
                 switch_spatial_to_temporal_mix: bool = False):
        super().__init__()
        self.merge_strategy = merge_strategy
        self.switch_spatial_to_temporal_mix = switch_spatial_to_temporal_mix
        if merge_strategy == "fixed":
            self.register_buffer("mix_factor", torch.Tensor([alpha]))
        else:
            self.register_parameter("mix_factor", nn.Parameter(torch.Tensor([alpha])))

    def get_alpha(self, image_only_indicator, ndims):
        if self.merge_strategy == "fixed":
            alpha = self.mix_factor
        elif self.merge_strategy == "learned":
            alpha = torch.sigmoid(self.mix_factor)
        else:  # learned_with_images
            alpha = torch.where(image_only_indicator.bool(),
                                torch.ones(1, 1, device=image_only_indicator.device),
                                torch.sigmoid(self.mix_factor)[..., None])
            if ndims == 5:
                alpha = alpha[:, None, :, None, None]
            elif ndims == 3:
                alpha = alpha.reshape(-1)[:, None, None]
        return alpha

    def forward(self, x_spatial, x_temporal, image_only_indicator=None):
        alpha = self.get_alpha(image_only_indicator, x_spatial.ndim).to(x_spatial.dtype)
        if self.switch_spatial_to_temporal_mix:
            alpha = 1.0 - alpha
        return alpha * x_spatial + (1.0 - alpha) * x_temporal


class SpatioTemporalResBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, eps: float = 1e-6,
                 merge_factor: float = 0.5, merge_strategy: str = "learned_with_images",
                 switch_spatial_to_temporal_mix: bool = False, temporal_eps: Optional[float] = None):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, out_channels, temb_channels,
                                                      temporal_eps if temporal_eps is not None else eps)
        self.time_mixer = AlphaBlender(merge_factor, merge_strategy, switch_spatial_to_temporal_mix)

    def forward(self, hidden_states, temb, image_only_indicator):
        num_frames = image_only_indicator.shape[-1]
        hidden_states = self.spatial_res_block(hidden_states, temb)
        bf, c, h, w = hidden_states.shape
        b = bf // num_frames
        hs_mix = hidden_states[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        hs = hidden_states[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        if temb is not None:
            temb = temb.reshape(b, num_frames, -1)
        hs = self.temporal_res_block(hs, temb)
        hs = self.time_mixer(x_spatial=hs_mix, x_temporal=hs, image_only_indicator=image_only_indicator)
        return hs.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# --------------------------------------------------------------------------------------------
# attention (diffusers models/attention.py, attention_processor.py)
# --------------------------------------------------------------------------------------------
class Attention(nn.Module):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None):  # AttnProcessor2_0
        b, l, _ = hidden_states.shape
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.to_q(hidden_states)
        k = self.to_k(ctx)
        v = self.to_v(ctx)
        hd = q.shape[-1] // self.heads
        q = q.view(b, -1, self.heads, hd).transpose(1, 2)
        k = k.view(b, -1, self.heads, hd).transpose(1, 2)
        v = v.view(b, -1, self.heads, hd).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, -1, self.heads * hd).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)  # exact (erf) GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0),
                                  nn.Linear(inner, dim_out if dim_out is not None else dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, head_dim: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, None, heads, head_dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, cross_attention_dim, heads, head_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        x = self.ff(self.norm3(x)) + x
        return x


class TemporalBasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, time_mix_inner_dim: int, heads: int, head_dim: int,
                 cross_attention_dim: Optional[int]):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim)
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = Attention(time_mix_inner_dim, None, heads, head_dim)
        if cross_attention_dim is not None:
            self.norm2 = nn.LayerNorm(time_mix_inner_dim)
            self.attn2 = Attention(time_mix_inner_dim, cross_attention_dim, heads, head_dim)
        else:
            self.norm2 = None
            self.attn2 = None
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = FeedForward(time_mix_inner_dim)

    def forward(self, hidden_states, num_frames, encoder_hidden_states=None):
        bf, s, c = hidden_states.shape
        b = bf // num_frames
        x = hidden_states[None, :].reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
        residual = x
        x = self.ff_in(self.norm_in(x))
        if self.is_res:
            x = x + residual
        x = self.attn1(self.norm1(x), None) + x
        if self.attn2 is not None:
            x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        ff = self.ff(self.norm3(x))
        x = ff + x if self.is_res else ff
        x = x[None, :].reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(b * num_frames, s, c)
        return x


class TransformerSpatioTemporalModel(nn.Module):
    """diffusers models/transformers/transformer_temporal.py (0.27.2).

    ``time_context_order``: "s_major" is the 0.27.2 behaviour the reference pins (the first-frame
    context is broadcast as [S, B] while the temporal block's rows are [B, S]); "b_major" is the
    later-release fix.  See SURVEY.md A.5.
    """

    def __init__(self, num_attention_heads: int, attention_head_dim: int, in_channels: int,
                 num_layers: int = 1, cross_attention_dim: Optional[int] = None,
                 time_context_order: str = "s_major"):
        super().__init__()
        inner = num_attention_heads * attention_head_dim
        self.time_context_order = time_context_order
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([
            BasicTransformerBlock(inner, num_attention_heads, attention_head_dim, cross_attention_dim)
            for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList([
            TemporalBasicTransformerBlock(inner, inner, num_attention_heads, attention_head_dim, cross_attention_dim)
            for _ in range(num_layers)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_proj = Timesteps(in_channels, True, 0)
        self.time_mixer = AlphaBlender(alpha=0.5, merge_strategy="learned_with_images")
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, hidden_states, encoder_hidden_states, image_only_indicator):
        bf, _, h, w = hidden_states.shape
        num_frames = image_only_indicator.shape[-1]
        b = bf // num_frames
        tc = encoder_hidden_states
        tc_first = tc[None, :].reshape(b, num_frames, -1, tc.shape[-1])[:, 0]
        if self.time_context_order == "s_major":
            tc = tc_first[None, :].broadcast_to(h * w, b, tc_first.shape[-2], tc.shape[-1])
            tc = tc.reshape(h * w * b, tc_first.shape[-2], tc.shape[-1])
        else:
            tc = tc_first[:, None].broadcast_to(b, h * w, tc_first.shape[-2], tc.shape[-1])
            tc = tc.reshape(b * h * w, tc_first.shape[-2], tc.shape[-1])
        residual = hidden_states
        x = self.norm(hidden_states)
        inner = x.shape[1]
        x = x.permute(0, 2, 3, 1).reshape(bf, h * w, inner)
        x = self.proj_in(x)
        frame_idx = torch.arange(num_frames, device=x.device).repeat(b, 1).reshape(-1)
        t_emb = self.time_proj(frame_idx).to(dtype=x.dtype)
        emb = self.time_pos_embed(t_emb)[:, None, :]
        for block, tblock in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            x = block(x, encoder_hidden_states)
            x_mix = x + emb
            x_mix = tblock(x_mix, num_frames=num_frames, encoder_hidden_states=tc)
            x = self.time_mixer(x_spatial=x, x_temporal=x_mix, image_only_indicator=image_only_indicator)
        x = self.proj_out(x)
        x = x.reshape(bf, h, w, inner).permute(0, 3, 1, 2).contiguous()
        return x + residual


# --------------------------------------------------------------------------------------------
# resampling (diffusers models/downsampling.py, upsampling.py)
# --------------------------------------------------------------------------------------------
class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# --------------------------------------------------------------------------------------------
# UNet blocks (diffusers models/unets/unet_3d_blocks.py)
# --------------------------------------------------------------------------------------------
class DownBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, **_):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=1e-5)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, image_only_indicator, encoder_hidden_states=None):
        out = ()
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, temb, image_only_indicator)
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class CrossAttnDownBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample,
                 num_attention_heads, cross_attention_dim, transformer_layers=1, time_context_order="s_major"):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=1e-6)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads,
                                           in_channels=out_channels, num_layers=transformer_layers,
                                           cross_attention_dim=cross_attention_dim,
                                           time_context_order=time_context_order)
            for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, image_only_indicator, encoder_hidden_states=None):
        out = ()
        for resnet, attn in zip(self.resnets, self.attentions):
            hidden_states = resnet(hidden_states, temb, image_only_indicator)
            hidden_states = attn(hidden_states, encoder_hidden_states, image_only_indicator)
            out += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            out += (hidden_states,)
        return hidden_states, out


class UNetMidBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, num_attention_heads, cross_attention_dim,
                 num_layers=1, transformer_layers=1, time_context_order="s_major"):
        super().__init__()
        resnets = [SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5)]
        attentions = []
        for _ in range(num_layers):
            attentions.append(TransformerSpatioTemporalModel(
                num_attention_heads, in_channels // num_attention_heads, in_channels=in_channels,
                num_layers=transformer_layers, cross_attention_dim=cross_attention_dim,
                time_context_order=time_context_order))
            resnets.append(SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps=1e-5))
        self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)

    def forward(self, hidden_states, temb, encoder_hidden_states, image_only_indicator):
        hidden_states = self.resnets[0](hidden_states, temb, image_only_indicator)
        for attn, resnet in zip(self.attentions, self.resnets[1:]):
            hidden_states = attn(hidden_states, encoder_hidden_states, image_only_indicator)
            hidden_states = resnet(hidden_states, temb, image_only_indicator)
        return hidden_states


class UpBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                 add_upsample, **_):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(rin + skip, out_channels, temb_channels, eps=1e-6))
        self.resnets = nn.ModuleList(resnets)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, image_only_indicator,
                encoder_hidden_states=None):
        for resnet in self.resnets:
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers,
                 add_upsample, num_attention_heads, cross_attention_dim, transformer_layers=1,
                 time_context_order="s_major"):
        super().__init__()
        resnets, attentions = [], []
        for i in range(num_layers):
            skip = in_channels if i == num_layers - 1 else out_channels
            rin = prev_output_channel if i == 0 else out_channels
            resnets.append(SpatioTemporalResBlock(rin + skip, out_channels, temb_channels, eps=1e-6))
            attentions.append(TransformerSpatioTemporalModel(
                num_attention_heads, out_channels // num_attention_heads, in_channels=out_channels,
                num_layers=transformer_layers, cross_attention_dim=cross_attention_dim,
                time_context_order=time_context_order))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList(attentions)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, image_only_indicator,
                encoder_hidden_states=None):
        for resnet, attn in zip(self.resnets, self.attentions):
            res = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, res], dim=1)
            hidden_states = resnet(hidden_states, temb, image_only_indicator)
            hidden_states = attn(hidden_states, encoder_hidden_states, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


_DOWN = {"CrossAttnDownBlockSpatioTemporal": CrossAttnDownBlockSpatioTemporal,
         "DownBlockSpatioTemporal": DownBlockSpatioTemporal}
_UP = {"CrossAttnUpBlockSpatioTemporal": CrossAttnUpBlockSpatioTemporal,
       "UpBlockSpatioTemporal": UpBlockSpatioTemporal}

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


class _Encoder(nn.Module):
    """Shared constructor of the embedding + down + mid part (both models build it the same way:
    controlnet.py:100-192 and the diffusers UNet __init__)."""

    def _build_encoder(self, cfg, time_context_order):
        boc = cfg["block_out_channels"]
        n = len(cfg["down_block_types"])
        heads = _tup(cfg["num_attention_heads"], n)
        xdim = _tup(cfg["cross_attention_dim"], n)
        lpb = _tup(cfg["layers_per_block"], n)
        tlpb = _tup(cfg["transformer_layers_per_block"], n)
        temb_dim = boc[0] * 4
        self.conv_in = nn.Conv2d(cfg["in_channels"], boc[0], 3, padding=1)
        self.time_proj = Timesteps(boc[0], True, 0)
        self.time_embedding = TimestepEmbedding(boc[0], temb_dim)
        self.add_time_proj = Timesteps(cfg["addition_time_embed_dim"], True, 0)
        self.add_embedding = TimestepEmbedding(cfg["projection_class_embeddings_input_dim"], temb_dim)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(cfg["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            kw = dict(in_channels=in_ch, out_channels=out_ch, temb_channels=temb_dim, num_layers=lpb[i],
                      add_downsample=(i != n - 1))
            if t.startswith("CrossAttn"):
                kw.update(num_attention_heads=heads[i], cross_attention_dim=xdim[i],
                          transformer_layers=tlpb[i], time_context_order=time_context_order)
            self.down_blocks.append(_DOWN[t](**kw))
        self.mid_block = UNetMidBlockSpatioTemporal(
            boc[-1], temb_dim, num_attention_heads=heads[-1], cross_attention_dim=xdim[-1],
            transformer_layers=tlpb[-1], time_context_order=time_context_order)
        return temb_dim, heads, xdim, lpb, tlpb

    def _embed(self, sample, timestep, added_time_ids):
        # controlnet.py:262-283 == unet_spatio_temporal_condition.py:64-85
        timesteps = timestep
        if len(timesteps.shape) == 0:
            timesteps = timesteps[None].to(sample.device)
        batch_size = sample.shape[0]
        timesteps = timesteps.expand(batch_size)
        t_emb = self.time_proj(timesteps).to(dtype=sample.dtype)
        emb = self.time_embedding(t_emb)
        time_embeds = self.add_time_proj(added_time_ids.flatten())
        time_embeds = time_embeds.reshape((batch_size, -1)).to(emb.dtype)
        return emb + self.add_embedding(time_embeds)

    def _down_mid(self, sample, emb, encoder_hidden_states, image_only_indicator):
        res = (sample,)
        for blk in self.down_blocks:
            sample, r = blk(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                            image_only_indicator=image_only_indicator)
            res += r
        return sample, res


class UNetSpatioTemporalConditionModel(_Encoder):
    """Restates unet_spatio_temporal_condition.py:13-171 on top of the diffusers constructor."""

    def __init__(self, time_context_order: str = "s_major", **overrides):
        super().__init__()
        cfg = dict(SVD_CONFIG); cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        temb_dim, heads, xdim, lpb, tlpb = self._build_encoder(cfg, time_context_order)
        boc = cfg["block_out_channels"]
        n = len(boc)
        rboc, rheads = list(reversed(boc)), list(reversed(heads))
        rlpb, rxdim, rtlpb = list(reversed(lpb)), list(reversed(xdim)), list(reversed(tlpb))
        self.up_blocks = nn.ModuleList()
        out_ch = rboc[0]
        for i, t in enumerate(cfg["up_block_types"]):
            prev, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            kw = dict(in_channels=in_ch, prev_output_channel=prev, out_channels=out_ch,
                      temb_channels=temb_dim, num_layers=rlpb[i] + 1, add_upsample=(i != n - 1))
            if t.startswith("CrossAttn"):
                kw.update(num_attention_heads=rheads[i], cross_attention_dim=rxdim[i],
                          transformer_layers=rtlpb[i], time_context_order=time_context_order)
            self.up_blocks.append(_UP[t](**kw))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], cfg["out_channels"], 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids,
                down_block_additional_residuals=None, mid_block_additional_residuals=None,
                return_dict: bool = True):
        is_controlnet = mid_block_additional_residuals is not None and down_block_additional_residuals is not None
        emb = self._embed(sample, timestep, added_time_ids)
        batch_size, num_frames = sample.shape[:2]
        sample = sample.flatten(0, 1)
        emb = emb.repeat_interleave(num_frames, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(num_frames, dim=0)
        sample = self.conv_in(sample)
        image_only_indicator = torch.zeros(batch_size, num_frames, dtype=sample.dtype, device=sample.device)
        sample, res = self._down_mid(sample, emb, encoder_hidden_states, image_only_indicator)
        if is_controlnet:
            res = tuple(r + a for r, a in zip(res, down_block_additional_residuals))
        sample = self.mid_block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                image_only_indicator=image_only_indicator)
        if is_controlnet:
            sample = sample + mid_block_additional_residuals
        for blk in self.up_blocks:
            k = len(blk.resnets)
            r, res = res[-k:], res[:-k]
            sample = blk(hidden_states=sample, temb=emb, res_hidden_states_tuple=r,
                         encoder_hidden_states=encoder_hidden_states,
                         image_only_indicator=image_only_indicator)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        sample = sample.reshape(batch_size, num_frames, *sample.shape[1:])
        if not return_dict:
            return (sample,)
        return SimpleNamespace(sample=sample)


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class ControlNetModel(_Encoder):
    """Restates src/ctrlv/models/controlnet.py:20-351."""

    def __init__(self, time_context_order: str = "s_major", **overrides):
        super().__init__()
        cfg = {k: v for k, v in SVD_CONFIG.items() if k not in ("out_channels", "up_block_types")}
        cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        boc = cfg["block_out_channels"]
        if len(boc) != len(cfg["down_block_types"]):  # controlnet.py:80-83
            raise ValueError("Must provide the same number of `block_out_channels` as `down_block_types`.")
        _, _, _, lpb, _ = self._build_encoder(cfg, time_context_order)
        self.control_conv_in = nn.Conv2d(cfg["in_channels"] // 2, boc[0], 3, padding=1)  # :136-141
        blocks = [zero_module(nn.Conv2d(boc[0], boc[0], 1))]  # :148-150
        for i, ch in enumerate(boc):
            for _ in range(lpb[i]):
                blocks.append(zero_module(nn.Conv2d(ch, ch, 1)))  # :172-175
            if i != len(boc) - 1:
                blocks.append(zero_module(nn.Conv2d(ch, ch, 1)))  # :177-180
        self.controlnet_down_blocks = nn.ModuleList(blocks)
        self.controlnet_mid_block = zero_module(nn.Conv2d(boc[-1], boc[-1], 1))  # :183-185

    @classmethod
    def from_unet(cls, unet: UNetSpatioTemporalConditionModel, load_weights_from_unet: bool = True):
        c = unet.config  # controlnet.py:197-224
        ctrl = cls(in_channels=c.in_channels, down_block_types=c.down_block_types,
                   block_out_channels=c.block_out_channels, addition_time_embed_dim=c.addition_time_embed_dim,
                   projection_class_embeddings_input_dim=c.projection_class_embeddings_input_dim,
                   layers_per_block=c.layers_per_block, cross_attention_dim=c.cross_attention_dim,
                   transformer_layers_per_block=c.transformer_layers_per_block,
                   num_attention_heads=c.num_attention_heads, num_frames=c.num_frames)
        if load_weights_from_unet:
            usd, csd = unet.state_dict(), ctrl.state_dict()
            with torch.no_grad():
                for k in csd:
                    if k in usd:
                        csd[k].copy_(usd[k])
        return ctrl

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, control_cond=None,
                conditioning_scale: float = 1.0, return_dict: bool = True):
        emb = self._embed(sample, timestep, added_time_ids)
        batch_size, num_frames = sample.shape[:2]
        sample = sample.flatten(0, 1)
        control_cond = control_cond.flatten(0, 1)
        emb = emb.repeat_interleave(num_frames, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(num_frames, dim=0)
        sample = self.conv_in(sample) + self.control_conv_in(control_cond)
        image_only_indicator = torch.zeros(batch_size, num_frames, dtype=sample.dtype, device=sample.device)
        sample, res = self._down_mid(sample, emb, encoder_hidden_states, image_only_indicator)
        sample = self.mid_block(hidden_states=sample, temb=emb, encoder_hidden_states=encoder_hidden_states,
                                image_only_indicator=image_only_indicator)
        res = [blk(r) * conditioning_scale for r, blk in zip(res, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(sample) * conditioning_scale
        if not return_dict:
            return (res, mid)
        return SimpleNamespace(down_block_res_samples=res, mid_block_res_sample=mid)


def randomize_zero_convs(ctrl: ControlNetModel, std: float = 0.02, seed: int = 1) -> None:
    """The zero-convs are zero at construction (controlnet.py:148-185), which would leave the
    injection path untested under random init; re-draw them N(0, std^2)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    with torch.no_grad():
        for m in list(ctrl.controlnet_down_blocks) + [ctrl.controlnet_mid_block]:
            m.weight.copy_(torch.randn(m.weight.shape, generator=g) * std)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * std)


# a reduced configuration for fast CPU/GPU parity tests (same topology, 64-multiple channels)
TINY_CONFIG = dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4),
                   cross_attention_dim=128, addition_time_embed_dim=32,
                   projection_class_embeddings_input_dim=96, num_frames=4)
