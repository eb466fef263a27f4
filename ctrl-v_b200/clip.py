"""CLIP image conditioning of the Box2Video pipelines on the sm_100a kernels.

  * `CLIPVisionModelWithProjection`: drop-in for the pipelines' `image_encoder`
    (pipeline_video_control.py:30; transformers' CLIP vision tower, ViT-H/14 for SVD) with the
    transformers state-dict key names; `model(pixel_values).image_embeds`.
  * `encode_image`: diffusers' `_encode_image` for a tensor image in [0, 1] (reached from
    pipeline_video_control.py:220): 2x-1, `_resize_with_antialiasing` to 224 (restated in
    src/ctrlv/bbox_generator_baseline/utils/image_encoder.py:184-290), back to [0,1], CLIP mean/std.

Runs once per clip.  Kernels: Gaussian blur + bicubic resize + patch gather (`csrc/clip.cu`), the
tcgen05 implicit GEMM for every projection (patch embedding = GEMM over gathered patches with the
position embedding as a per-row bias; GELU through the GEGLU epilogue with a constant-one value
column), LayerNorm, and per-head attention as two GEMMs around the row softmax (head_dim 80 is
zero-padded to 128 in the packed weights).  No PyTorch math on the path.
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional

import torch

from . import ops
from .models import BF16, _f, _w

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)
CLIP_VIT_H_CONFIG = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                         image_size=224, patch_size=14, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)


def param_spec(cfg: dict) -> "OrderedDict[str, tuple]":
    d, inner, p = cfg["hidden_size"], cfg["intermediate_size"], cfg["patch_size"]
    n = (cfg["image_size"] // p) ** 2 + 1
    spec: "OrderedDict[str, tuple]" = OrderedDict()
    v = "vision_model."
    spec[v + "embeddings.class_embedding"] = (d,)
    spec[v + "embeddings.patch_embedding.weight"] = (d, 3, p, p)
    spec[v + "embeddings.position_embedding.weight"] = (n, d)
    for nm in ("pre_layrnorm", "post_layernorm"):
        spec[v + nm + ".weight"] = (d,); spec[v + nm + ".bias"] = (d,)
    for i in range(cfg["num_hidden_layers"]):
        l = f"{v}encoder.layers.{i}."
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            spec[l + f"self_attn.{nm}.weight"] = (d, d); spec[l + f"self_attn.{nm}.bias"] = (d,)
        for nm in ("layer_norm1", "layer_norm2"):
            spec[l + nm + ".weight"] = (d,); spec[l + nm + ".bias"] = (d,)
        spec[l + "mlp.fc1.weight"] = (inner, d); spec[l + "mlp.fc1.bias"] = (inner,)
        spec[l + "mlp.fc2.weight"] = (d, inner); spec[l + "mlp.fc2.bias"] = (d,)
    spec["visual_projection.weight"] = (cfg["projection_dim"], d)
    return spec


def random_state_dict(cfg: dict, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator("cpu").manual_seed(seed)
    sd = OrderedDict()
    for k, shape in param_spec(cfg).items():
        if "norm" in k:
            sd[k] = torch.ones(shape) if k.endswith("weight") else torch.zeros(shape)
        elif k.endswith("bias"):
            sd[k] = 0.02 * torch.randn(shape, generator=g)
        elif k.endswith("embedding") or "position_embedding" in k:
            sd[k] = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_in = 1
            for dd in shape[1:]:
                fan_in *= dd
            sd[k] = torch.randn(shape, generator=g) / fan_in ** 0.5
    return sd


class _Layer:
    def __init__(self, sd, pfx, d, heads, dp):
        hd = d // heads
        self.ln1 = (_f(sd[pfx + "layer_norm1.weight"]), _f(sd[pfx + "layer_norm1.bias"]))
        self.ln2 = (_f(sd[pfx + "layer_norm2.weight"]), _f(sd[pfx + "layer_norm2.bias"]))

        def per_head(name):  # [d, d] -> [heads][dp, d] with zero rows past head_dim, bias [heads][dp]
            w = sd[pfx + f"self_attn.{name}.weight"].float().reshape(heads, hd, d)
            b = sd[pfx + f"self_attn.{name}.bias"].float().reshape(heads, hd)
            wp = torch.zeros((heads, dp, d), dtype=torch.float32, device=w.device); wp[:, :hd] = w
            bp = torch.zeros((heads, dp), dtype=torch.float32, device=w.device); bp[:, :hd] = b
            return wp, bp

        wq, bq = per_head("q_proj")
        wk, bk = per_head("k_proj")
        wv, bv = per_head("v_proj")
        self.wq, self.bq = _w(wq.reshape(heads * dp, d)), _f(bq.reshape(-1))   # all heads in one GEMM
        self.wk = [_w(wk[h]) for h in range(heads)]
        self.bk = [_f(bk[h]) for h in range(heads)]
        self.wv = [_w(wv[h]) for h in range(heads)]
        self.bv = [_f(bv[h]) for h in range(heads)]
        wo = sd[pfx + "self_attn.out_proj.weight"].float().reshape(d, heads, hd)
        wop = torch.zeros((d, heads, dp), dtype=torch.float32, device=wo.device); wop[:, :, :hd] = wo
        self.wo, self.bo = _w(wop.reshape(d, heads * dp)), _f(sd[pfx + "self_attn.out_proj.bias"])
        # fc1 + GELU through the GEGLU epilogue: (value, gate) column pairs with value == 1
        w1, b1 = sd[pfx + "mlp.fc1.weight"].float(), sd[pfx + "mlp.fc1.bias"].float()
        inner = w1.shape[0]
        wi = torch.zeros((2 * inner, d), dtype=torch.float32, device=w1.device); wi[1::2] = w1
        bi = torch.ones((2 * inner,), dtype=torch.float32, device=w1.device); bi[1::2] = b1
        self.w1, self.b1 = _w(wi), _f(bi)
        self.w2, self.b2 = _w(sd[pfx + "mlp.fc2.weight"].float()), _f(sd[pfx + "mlp.fc2.bias"])


class CLIPVisionModelWithProjection(torch.nn.Module):
    def __init__(self, state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0, **overrides):
        super().__init__()
        cfg = dict(CLIP_VIT_H_CONFIG)
        cfg.update(overrides)
        if cfg["hidden_act"] != "gelu":
            raise NotImplementedError("only hidden_act='gelu' (CLIP ViT-H, the SVD image encoder) is built")
        if cfg["hidden_size"] % 64 or cfg["intermediate_size"] % 64 or cfg["projection_dim"] % 32:
            raise ValueError("hidden/intermediate sizes must be multiples of 64 and projection_dim of 32")
        self.cfg = cfg
        self.config = SimpleNamespace(**cfg)
        self.dtype = BF16
        self._sd: Dict[str, torch.Tensor] = {}
        self.load_state_dict(state_dict if state_dict is not None else random_state_dict(cfg, seed))

    def state_dict(self, *a, **k):
        return OrderedDict(self._sd)

    def load_state_dict(self, sd, strict: bool = True):
        sd = {k: v for k, v in sd.items() if not k.endswith("position_ids")}
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
        """`<path>/<subfolder>/{config.json, model[.variant].safetensors | pytorch_model.bin}` (transformers layout)."""
        import json
        import os
        from . import checkpoint
        kwargs.pop("torch_dtype", None)
        d = os.path.join(path, subfolder) if subfolder else path
        with open(os.path.join(d, "config.json"), "r", encoding="utf-8") as f:
            config = json.load(f)
        stem = "model" + (f".{variant}" if variant else "")
        if os.path.isfile(os.path.join(d, stem + ".safetensors")):
            sd = checkpoint.read_safetensors(os.path.join(d, stem + ".safetensors"))
        elif os.path.isfile(os.path.join(d, "pytorch_model.bin")):
            sd = torch.load(os.path.join(d, "pytorch_model.bin"), map_location="cpu", weights_only=True)
        else:
            raise OSError(f"no {stem}.safetensors / pytorch_model.bin in {d}")
        over = {k: v for k, v in config.items() if k in CLIP_VIT_H_CONFIG}
        over.update(kwargs)
        return cls(state_dict=sd, **over)

    def save_pretrained(self, save_directory: str, subfolder: Optional[str] = None, variant: Optional[str] = None):
        """transformers layout: config.json + model[.variant].safetensors."""
        import json
        import os
        from . import checkpoint
        d = os.path.join(save_directory, subfolder) if subfolder else save_directory
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "config.json"), "w", encoding="utf-8") as f:
            json.dump(dict(architectures=["CLIPVisionModelWithProjection"], model_type="clip_vision_model", **self.cfg), f, indent=2)
        checkpoint.write_safetensors(os.path.join(d, "model" + (f".{variant}" if variant else "") + ".safetensors"),
                                     self._sd, {"format": "pt"})

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def _pack(self):
        sd, cfg = self._sd, self.cfg
        d, heads, p = cfg["hidden_size"], cfg["num_attention_heads"], cfg["patch_size"]
        self.hd = d // heads
        self.dp = (self.hd + 63) // 64 * 64  # per-head width padded to the GEMM's K granularity
        v = "vision_model."
        k = 3 * p * p
        self.kpad = (k + 63) // 64 * 64
        wp = torch.zeros((d, self.kpad), dtype=torch.float32, device=sd[v + "embeddings.patch_embedding.weight"].device)
        wp[:, :k] = sd[v + "embeddings.patch_embedding.weight"].float().reshape(d, k)
        self.w_patch = _w(wp)
        pos = sd[v + "embeddings.position_embedding.weight"].float()
        self.pos_patches = _f(pos[1:])
        self.cls_row = (sd[v + "embeddings.class_embedding"].float() + pos[0]).to("cuda", BF16)
        self.pre_ln = (_f(sd[v + "pre_layrnorm.weight"]), _f(sd[v + "pre_layrnorm.bias"]))
        self.post_ln = (_f(sd[v + "post_layernorm.weight"]), _f(sd[v + "post_layernorm.bias"]))
        self.layers = [_Layer(sd, f"{v}encoder.layers.{i}.", d, heads, self.dp) for i in range(cfg["num_hidden_layers"])]
        self.w_proj = _w(sd["visual_projection.weight"].float())
        self.mean = torch.tensor(CLIP_MEAN, dtype=torch.float32, device="cuda")
        self.std = torch.tensor(CLIP_STD, dtype=torch.float32, device="cuda")
        self.zero3 = torch.zeros(3, dtype=torch.float32, device="cuda")
        self.one3 = torch.ones(3, dtype=torch.float32, device="cuda")

    # ---- forward ------------------------------------------------------------------------------------
    def _tower(self, rows: torch.Tensor, B: int) -> torch.Tensor:
        """rows: gathered, normalised patches [B*G, kpad] bf16 -> image_embeds [B, projection_dim] fp32."""
        cfg = self.cfg
        d, heads, eps = cfg["hidden_size"], cfg["num_attention_heads"], cfg["layer_norm_eps"]
        G = (cfg["image_size"] // cfg["patch_size"]) ** 2
        S = G + 1
        Sp = (S + 63) // 64 * 64
        dp = self.dp
        x = torch.empty((B * S, d), dtype=BF16, device="cuda")
        for b in range(B):
            x[b * S].copy_(self.cls_row)
            ops.linear(rows[b * G:(b + 1) * G], self.w_patch, rowbias=self.pos_patches, rb_mode=2, rb_div=1, rb_mod=G,
                       out=x[b * S + 1:(b + 1) * S])
        x = ops.layernorm(x, self.pre_ln[0], self.pre_ln[1], eps)
        scale = float(self.hd) ** -0.5
        a_pad = torch.zeros((Sp, d), dtype=BF16, device="cuda")           # one image's LN1 output, zero rows past S
        kh = torch.zeros((Sp, dp), dtype=BF16, device="cuda")
        vt = torch.empty((dp, Sp), dtype=BF16, device="cuda")
        scores = torch.empty((S, Sp), dtype=torch.float32, device="cuda")
        probs = torch.zeros((S, Sp), dtype=BF16, device="cuda")
        o = torch.empty((B * S, heads * dp), dtype=BF16, device="cuda")
        for L in self.layers:
            a = ops.layernorm(x, L.ln1[0], L.ln1[1], eps)
            q = ops.linear(a, L.wq, bias=L.bq)                              # [B*S, heads*dp]
            for b in range(B):
                rs = slice(b * S, (b + 1) * S)
                a_pad[:S].copy_(a[rs])
                for h in range(heads):
                    ops.linear(a[rs], L.wk[h], bias=L.bk[h], out=kh[:S])    # K_h [S, dp] (rows past S stay zero)
                    ops.linear(L.wv[h], a_pad, out=vt)                      # V_h^T [dp, Sp] without bias
                    ops.linear(q[rs, h * dp:(h + 1) * dp], kh, out_f32=scores)
                    ops.softmax_rows(scores[:, :S], scale, out=probs[:, :S])
                    ops.linear(probs, vt, bias=L.bv[h], out=o[rs, h * dp:(h + 1) * dp])
            x = ops.linear(o, L.wo, bias=L.bo, res1=x)
            a2 = ops.layernorm(x, L.ln2[0], L.ln2[1], eps)
            hmid = ops.linear(a2, L.w1, bias=L.b1, geglu=True)              # 1 * gelu(fc1)
            x = ops.linear(hmid, L.w2, bias=L.b2, res1=x)
        pooled = torch.empty((B, d), dtype=BF16, device="cuda")
        for b in range(B):
            pooled[b].copy_(x[b * S])
        pooled = ops.layernorm(pooled, self.post_ln[0], self.post_ln[1], eps)
        out = torch.empty((B, cfg["projection_dim"]), dtype=torch.float32, device="cuda")
        ops.linear(pooled, self.w_proj, out_f32=out)
        return out

    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, return_dict: bool = True):
        """pixel_values [B, 3, S, S]: CLIP-normalised pixels (what `feature_extractor` returns)."""
        s = self.cfg["image_size"]
        if pixel_values.ndim != 4 or tuple(pixel_values.shape[1:]) != (3, s, s):
            raise ValueError(f"pixel_values must be [B, 3, {s}, {s}], got {tuple(pixel_values.shape)}")
        img = pixel_values.to("cuda", torch.float32).contiguous()
        rows = ops.clip_patchify(img, self.cfg["patch_size"], self.zero3, self.one3)
        emb = self._tower(rows, img.shape[0])
        if not return_dict:
            return (emb,)
        return SimpleNamespace(image_embeds=emb)

    @torch.no_grad()
    def encode_image(self, image01: torch.Tensor, clamp: bool = False) -> torch.Tensor:
        """`_encode_image` for a [B, 3, H, W] tensor in [0, 1] -> image embeddings [B, 1, D] fp32."""
        if image01.ndim != 4 or image01.shape[1] != 3:
            raise ValueError(f"image must be [B, 3, H, W] in [0, 1], got {tuple(image01.shape)}")
        s = self.cfg["image_size"]
        img = image01.to("cuda", torch.float32).contiguous()
        # The reference maps to [-1, 1], resizes, and maps back.  Both filters have weights that sum to
        # one, so they commute with that affine map: resize the [0, 1] image directly.
        small = resize_with_antialiasing(img, (s, s))
        rows = ops.clip_patchify(small, self.cfg["patch_size"], self.mean, self.std, clamp01=clamp)
        return self._tower(rows, img.shape[0]).unsqueeze(1)


def _blur_params(n_in: int, n_out: int):
    factor = n_in / n_out
    sigma = max((factor - 1.0) / 2.0, 0.001)
    ks = int(max(2.0 * 2 * sigma, 3))
    return sigma, ks + (1 - ks % 2)


def _taps(ks: int, sigma: float) -> torch.Tensor:  # host-side constants of the filter
    x = torch.arange(ks, dtype=torch.float32) - ks // 2
    g = torch.exp(-x.pow(2.0) / (2 * torch.tensor(sigma, dtype=torch.float32).pow(2.0)))
    return (g / g.sum()).to("cuda")


def resize_with_antialiasing(img: torch.Tensor, size) -> torch.Tensor:
    """`_resize_with_antialiasing(img, size)` on the device: separable Gaussian (reflect padding), then
    bicubic with align_corners=True; img [B, C, H, W] fp32."""
    H, W = img.shape[-2:]
    sy, ky = _blur_params(H, size[0])
    sx, kx = _blur_params(W, size[1])
    x = ops.blur1d_reflect(img, _taps(kx, sx), 0)
    x = ops.blur1d_reflect(x, _taps(ky, sy), 1)
    return ops.resize_bicubic_ac(x, size[0], size[1])
