"""ORACLE (test infrastructure — never imported by the product path).

Image-conditioning prologue of the reference pipelines:
  * `resize_with_antialiasing` restates `_resize_with_antialiasing`
    (/root/reference/src/ctrlv/bbox_generator_baseline/utils/image_encoder.py:184-290, called by
    src/ctrlv/utils/util.py:97-125 and by diffusers' `_encode_image` behind
    pipeline_video_control.py:220).  PINNED: tests/golden/resize_antialias.pt holds outputs of the
    reference's own functions (tests/golden/make_resize_golden.py).
  * `CLIPVisionModelWithProjection` restates transformers' modeling_clip vision tower
    (`image_encoder` of the pipelines, pipeline_video_control.py:30; transformers==4.45.2 pinned by the
    reference's requirements.txt:25, 5.5.0 installed here).  PINNED against the installed
    transformers implementation in tests/test_oracle.py.
  * `encode_image` restates diffusers' `StableVideoDiffusionPipeline._encode_image` for tensor input
    in [0, 1] (normalise to [-1,1], antialiased resize to 224, back to [0,1], CLIP mean/std).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
CLIP_VIT_H_CONFIG = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                         image_size=224, patch_size=14, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)
TINY_CLIP_CONFIG = dict(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                        image_size=28, patch_size=14, projection_dim=64, hidden_act="gelu", layer_norm_eps=1e-5)


def blur_params(n_in: int, n_out: int):
    """sigma and (odd) tap count for one axis: image_encoder.py:188-207."""
    factor = n_in / n_out
    sigma = max((factor - 1.0) / 2.0, 0.001)
    ks = int(max(2.0 * 2 * sigma, 3))
    if ks % 2 == 0:
        ks += 1
    return sigma, ks


def gaussian_taps(ks: int, sigma: float) -> torch.Tensor:
    """Normalised Gaussian window centred on ks // 2 (image_encoder.py:258-272)."""
    x = torch.arange(ks, dtype=torch.float32) - ks // 2
    g = torch.exp(-x.pow(2.0) / (2 * torch.tensor(sigma, dtype=torch.float32).pow(2.0)))
    return g / g.sum()


def _blur_axis(x: torch.Tensor, taps: torch.Tensor, dim: int) -> torch.Tensor:
    ks = taps.numel()
    front = (ks - 1) // 2
    rear = ks - 1 - front
    pad = (front, rear, 0, 0) if dim == -1 else (0, 0, front, rear)
    xp = F.pad(x, pad, mode="reflect")
    win = xp.unfold(dim, ks, 1)  # [..., n, ks]
    return (win * taps.to(x)).sum(-1)


def resize_with_antialiasing(x: torch.Tensor, size) -> torch.Tensor:
    h, w = x.shape[-2:]
    sy, ky = blur_params(h, size[0])
    sx, kx = blur_params(w, size[1])
    x = _blur_axis(x, gaussian_taps(kx, sx), -1)
    x = _blur_axis(x, gaussian_taps(ky, sy), -2)
    return F.interpolate(x, size=tuple(size), mode="bicubic", align_corners=True)


# --------------------------------------------------------------------------------------------
# CLIP vision tower (transformers models/clip/modeling_clip.py)
# --------------------------------------------------------------------------------------------
class _Attn(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.heads = heads
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))

    def forward(self, x):
        b, n, d = x.shape
        hd = d // self.heads
        q, k, v = (p(x).view(b, n, self.heads, hd).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        att = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, dim=-1)
        return self.out_proj((att @ v).transpose(1, 2).reshape(b, n, d))


class _MLP(nn.Module):
    def __init__(self, d, inner, act):
        super().__init__()
        self.fc1, self.fc2, self.act = nn.Linear(d, inner), nn.Linear(inner, d), act

    def forward(self, x):
        h = self.fc1(x)
        h = F.gelu(h) if self.act == "gelu" else h * torch.sigmoid(1.702 * h)  # "quick_gelu"
        return self.fc2(h)


class _Layer(nn.Module):
    def __init__(self, d, heads, inner, act, eps):
        super().__init__()
        self.self_attn = _Attn(d, heads)
        self.layer_norm1 = nn.LayerNorm(d, eps=eps)
        self.mlp = _MLP(d, inner, act)
        self.layer_norm2 = nn.LayerNorm(d, eps=eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, d, image_size, patch):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(d))
        self.patch_embedding = nn.Conv2d(3, d, patch, stride=patch, bias=False)
        n = (image_size // patch) ** 2 + 1
        self.position_embedding = nn.Embedding(n, d)
        self.register_buffer("position_ids", torch.arange(n)[None], persistent=False)

    def forward(self, pixel_values):
        p = self.patch_embedding(pixel_values).flatten(2).transpose(1, 2)
        cls = self.class_embedding.expand(p.shape[0], 1, -1)
        return torch.cat([cls, p], dim=1) + self.position_embedding(self.position_ids)


class _Encoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(cfg["hidden_size"], cfg["num_attention_heads"], cfg["intermediate_size"],
                                            cfg["hidden_act"], cfg["layer_norm_eps"])
                                     for _ in range(cfg["num_hidden_layers"])])


class _VisionTransformer(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        d, eps = cfg["hidden_size"], cfg["layer_norm_eps"]
        self.embeddings = _Embeddings(d, cfg["image_size"], cfg["patch_size"])
        self.pre_layrnorm = nn.LayerNorm(d, eps=eps)  # (sic) the key name of the checkpoints
        self.encoder = _Encoder(cfg)
        self.post_layernorm = nn.LayerNorm(d, eps=eps)

    def forward(self, pixel_values):
        x = self.pre_layrnorm(self.embeddings(pixel_values))
        for layer in self.encoder.layers:
            x = layer(x)
        return self.post_layernorm(x[:, 0])


class CLIPVisionModelWithProjection(nn.Module):
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(CLIP_VIT_H_CONFIG)
        cfg.update(overrides)
        self.cfg = cfg
        self.vision_model = _VisionTransformer(cfg)
        self.visual_projection = nn.Linear(cfg["hidden_size"], cfg["projection_dim"], bias=False)

    def forward(self, pixel_values):
        """-> image_embeds [B, projection_dim]"""
        return self.visual_projection(self.vision_model(pixel_values))


def clip_normalize(x01: torch.Tensor) -> torch.Tensor:
    mean = torch.tensor(CLIP_MEAN, dtype=x01.dtype, device=x01.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=x01.dtype, device=x01.device).view(1, 3, 1, 1)
    return (x01 - mean) / std


def encode_image(image_encoder: CLIPVisionModelWithProjection, image01: torch.Tensor, size=None, clamp: bool = False):
    """image01 [B, 3, H, W] in [0, 1] -> image embeddings [B, 1, D]: `_encode_image` (2x-1, antialiased
    resize, (x+1)/2, CLIP normalisation; `clamp` = the variant of src/ctrlv/utils/util.py:107-110)."""
    s = size or image_encoder.cfg["image_size"]
    x = resize_with_antialiasing(image01 * 2.0 - 1.0, (s, s))
    x = (x + 1.0) / 2.0
    if clamp:
        x = x.clamp(0.0, 1.0)
    return image_encoder(clip_normalize(x)).unsqueeze(1)
