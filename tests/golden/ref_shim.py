"""Runs the REFERENCE'S OWN in-repo code of the hot path in this container (test infrastructure).

The reference (`/root/reference/src/ctrlv`) is Python on top of `diffusers==0.27.2`, which is not
installed and not installable offline, so `import ctrlv.models` fails at its first line.  This module
installs a stand-in `diffusers` module tree in `sys.modules` whose block classes ARE the oracle's
restatements (`oracle/svd_oracle.py`, `oracle/sampling.py`) and whose pipeline base class is a small
test double, then loads the reference source files *from where they lie* (nothing is copied):

  src/ctrlv/models/unet_spatio_temporal_condition.py   (forward :31-171)
  src/ctrlv/models/controlnet.py                       (__init__ :53-195, from_unet :197-224, forward :226-351)
  src/ctrlv/pipelines/pipeline_video_control.py        (_encode_vae_condition :71-101, __call__ :105-360)
  src/ctrlv/pipelines/pipeline_video_diffusion.py      (bbox-predictor stage, __call__ :57-310)

What this pins: the oracle's restatement of those in-repo files (constructor layout and state-dict
keys, embedding path, conv_in + control_conv_in, residual wiring, zero-conv order, conditioning
scale, CFG duplication / combine, per-frame guidance, conditioning-frame overwrite, the order of
scheduler calls) against the reference code itself, bit for bit.  What it does NOT pin: the
arithmetic inside the diffusers blocks and scheduler, which both sides take from the oracle — that
part of the parity stays unpinned.

`/root/reference` exists only in the build container: callers skip when it is absent; the vectors
generated with this module are committed (`make_ref_golden.py` -> `ref_forward.pt`).
"""
from __future__ import annotations

import contextlib
import functools
import importlib.util
import inspect
import logging as _pylogging
import os
import sys
import types
from dataclasses import dataclass
from types import SimpleNamespace

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import sampling as S  # noqa: E402
from oracle import svd_oracle as O  # noqa: E402

REFERENCE_SRC = os.environ.get("CTRLV_REFERENCE_SRC", "/root/reference/src/ctrlv")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "models", "controlnet.py"))


# ------------------------------------------------------------------------------------------------
# stand-ins for the diffusers infrastructure classes (no arithmetic in here)
# ------------------------------------------------------------------------------------------------
class ConfigMixin:
    def __init__(self):
        pass


def register_to_config(init):
    """diffusers' decorator: records the constructor arguments (defaults included) as `.config`."""
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        params = [p for p in inspect.signature(init).parameters.values() if p.name != "self"]
        cfg = {p.name: p.default for p in params}
        cfg.update({p.name: a for p, a in zip(params, args)})
        cfg.update(kwargs)
        object.__setattr__(self, "config", SimpleNamespace(**cfg))
        init(self, *args, **kwargs)
    return inner


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device


class _EmptyMixin:
    def __init__(self):
        pass


class ControlNetModelOriginal(ModelMixin, ConfigMixin, _EmptyMixin):
    """Base class name only: the reference's constructor bypasses it (controlnet.py:72-75)."""


@dataclass
class ControlNetOutput:
    down_block_res_samples: tuple
    mid_block_res_sample: torch.Tensor


@dataclass
class UNetSpatioTemporalConditionOutput:
    sample: torch.Tensor = None


@dataclass
class StableVideoDiffusionPipelineOutput:
    frames: object = None


def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                   num_attention_heads=None, cross_attention_dim=None, transformer_layers_per_block=1, **_unused):
    """diffusers' factory: for the two SpatioTemporal block types it forwards only these arguments
    (resnet_eps / resnet_act_fn are not passed on)."""
    if down_block_type == "DownBlockSpatioTemporal":
        return O.DownBlockSpatioTemporal(in_channels=in_channels, out_channels=out_channels,
                                         temb_channels=temb_channels, num_layers=num_layers,
                                         add_downsample=add_downsample)
    if down_block_type == "CrossAttnDownBlockSpatioTemporal":
        return O.CrossAttnDownBlockSpatioTemporal(in_channels=in_channels, out_channels=out_channels,
                                                  temb_channels=temb_channels, num_layers=num_layers,
                                                  add_downsample=add_downsample,
                                                  num_attention_heads=num_attention_heads,
                                                  cross_attention_dim=cross_attention_dim,
                                                  transformer_layers=transformer_layers_per_block)
    raise ValueError(f"{down_block_type} does not exist.")


def _mid_block(in_channels, temb_channels, transformer_layers_per_block=1, cross_attention_dim=None,
               num_attention_heads=None, num_layers=1):
    return O.UNetMidBlockSpatioTemporal(in_channels, temb_channels, num_attention_heads=num_attention_heads,
                                        cross_attention_dim=cross_attention_dim, num_layers=num_layers,
                                        transformer_layers=transformer_layers_per_block)


class SchedulerShim:
    """diffusers EulerDiscreteScheduler surface over the oracle scheduler (`sampling.py`)."""
    order = 1

    def __init__(self):
        self._s = S.EulerDiscreteSchedulerOracle()

    def set_timesteps(self, num_inference_steps, device=None):
        self._s.set_timesteps(num_inference_steps)

    timesteps = property(lambda self: self._s.timesteps)
    sigmas = property(lambda self: self._s.sigmas)
    init_noise_sigma = property(lambda self: self._s.init_noise_sigma)

    def scale_model_input(self, sample, t):
        return self._s.scale_model_input(sample, t)

    def step(self, model_output, t, sample):
        return SimpleNamespace(prev_sample=self._s.step(model_output, t, sample))


class FakeVAE(nn.Module):
    """Test double for AutoencoderKLTemporalDecoder.encode(...).latent_dist.mode(): 8x average pool
    and a fixed 3 -> 4 channel mix.  Deterministic, so both sides of a comparison can call it."""
    config = SimpleNamespace(block_out_channels=(1, 1, 1, 1), force_upcast=False, scaling_factor=0.18215)
    dtype = torch.float32
    device = torch.device("cpu")

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(77)
        self.register_buffer("mix", torch.randn(4, 3, generator=g))

    def latents(self, x):
        p = torch.nn.functional.avg_pool2d(x.float(), 8)
        return torch.einsum("oc,bchw->bohw", self.mix, p)

    def encode(self, x):
        z = self.latents(x)
        return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: z))


class FakeImageEncoder:
    """Test double for CLIPVisionModelWithProjection: a fixed projection of the mean colour."""

    def __init__(self, dim):
        g = torch.Generator().manual_seed(78)
        self.w = torch.randn(3, dim, generator=g)

    def embeds(self, image01):
        return (image01.float().mean(dim=(2, 3)) @ self.w).unsqueeze(1)  # [B, 1, dim]


class VaeImageProcessor:
    def __init__(self, vae_scale_factor=8):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        assert isinstance(image, torch.Tensor) and image.shape[-2:] == (height, width)
        return 2.0 * image - 1.0  # [0, 1] -> [-1, 1]; no resize needed at the target size


def randn_tensor(shape, generator=None, device=None, dtype=None):
    return torch.randn(tuple(shape), generator=generator, dtype=dtype)


def _append_dims(x, target_dims):
    return x[(...,) + (None,) * (target_dims - x.ndim)]


class DiffusionPipeline:
    def __init__(self):
        pass

    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    _execution_device = torch.device("cpu")

    @contextlib.contextmanager
    def progress_bar(self, total=None):
        yield SimpleNamespace(update=lambda: None)

    def maybe_free_model_hooks(self):
        pass


class StableVideoDiffusionPipelineOriginal(DiffusionPipeline):
    """Test double for the diffusers base pipeline: only the helper methods the reference's
    `__call__`s use, stated from the published 0.27.2 pipeline; the encoders are
    the fakes above."""

    def __init__(self, vae, image_encoder, unet, scheduler, feature_extractor):
        DiffusionPipeline.__init__(self)
        self.register_modules(vae=vae, image_encoder=image_encoder, unet=unet, scheduler=scheduler,
                              feature_extractor=feature_extractor)
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.image_processor = VaeImageProcessor(vae_scale_factor=self.vae_scale_factor)

    guidance_scale = property(lambda self: self._guidance_scale)

    @property
    def do_classifier_free_guidance(self):
        if isinstance(self.guidance_scale, (int, float)):
            return self.guidance_scale > 1
        return self.guidance_scale.max() > 1

    def check_inputs(self, image, height, width):
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _encode_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        e = self.image_encoder.embeds(image).repeat(num_videos_per_prompt, 1, 1)
        return torch.cat([torch.zeros_like(e), e]) if do_classifier_free_guidance else e

    def _encode_vae_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        z = self.vae.encode(image).latent_dist.mode()
        if do_classifier_free_guidance:
            z = torch.cat([torch.zeros_like(z), z])
        return z.repeat(num_videos_per_prompt, 1, 1, 1)

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, dtype, batch_size,
                          num_videos_per_prompt, do_classifier_free_guidance):
        ids = [fps, motion_bucket_id, noise_aug_strength]
        passed = self.unet.config.addition_time_embed_dim * len(ids)
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(f"Model expects an added time embedding vector of length {expected}, but {passed} was created.")
        t = torch.tensor([ids], dtype=dtype).repeat(batch_size * num_videos_per_prompt, 1)
        return torch.cat([t, t]) if do_classifier_free_guidance else t

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device,
                        generator, latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        latents = randn_tensor(shape, generator=generator, dtype=dtype) if latents is None else latents.to(device)
        return latents * self.scheduler.init_noise_sigma


# ------------------------------------------------------------------------------------------------
# sys.modules plumbing
# ------------------------------------------------------------------------------------------------
_SHIMMED = ("diffusers", "ctrlv")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []  # behaves as a package for `from a.b import c`
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(modname, relpath, package):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REFERENCE_SRC, relpath))
    m = importlib.util.module_from_spec(spec)
    m.__package__ = package
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


def load_reference():
    """Returns a namespace with the reference's own classes: UNetSpatioTemporalConditionModel,
    ControlNetModel, StableVideoControlPipeline, VideoDiffusionPipeline."""
    if not available():
        raise FileNotFoundError(REFERENCE_SRC)
    if any(k in sys.modules and not getattr(sys.modules[k], "_ctrlv_ref_shim", False) for k in _SHIMMED):
        raise RuntimeError("a real `diffusers` / `ctrlv` is already imported; the shim must not shadow it")
    logging = SimpleNamespace(get_logger=_pylogging.getLogger)
    ident = lambda *_a, **_k: (lambda f: f)
    base = _mod("diffusers", UNetSpatioTemporalConditionModel=O.UNetSpatioTemporalConditionModel,
                StableVideoDiffusionPipeline=StableVideoDiffusionPipelineOriginal,
                EulerDiscreteScheduler=SchedulerShim, _ctrlv_ref_shim=True)
    _mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _mod("diffusers.models", ControlNetModel=ControlNetModelOriginal, AutoencoderKLTemporalDecoder=FakeVAE)
    _mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _mod("diffusers.models.unets")
    _mod("diffusers.models.unets.unet_3d_blocks", UNetMidBlockSpatioTemporal=_mid_block,
         get_down_block=get_down_block)
    _mod("diffusers.models.unets.unet_spatio_temporal_condition",
         UNetSpatioTemporalConditionOutput=UNetSpatioTemporalConditionOutput)
    _mod("diffusers.loaders", FromOriginalControlNetMixin=_EmptyMixin, PeftAdapterMixin=type("PeftAdapterMixin", (), {}))
    _mod("diffusers.models.embeddings", TimestepEmbedding=O.TimestepEmbedding, Timesteps=O.Timesteps)
    _mod("diffusers.utils", logging=logging, replace_example_docstring=ident, BaseOutput=object)
    _mod("diffusers.utils.torch_utils", randn_tensor=randn_tensor, is_compiled_module=lambda m: False)
    _mod("diffusers.models.controlnet", ControlNetOutput=ControlNetOutput, zero_module=O.zero_module)
    _mod("diffusers.pipelines")
    _mod("diffusers.pipelines.pipeline_utils", DiffusionPipeline=DiffusionPipeline)
    _mod("diffusers.pipelines.stable_video_diffusion")
    _mod("diffusers.pipelines.stable_video_diffusion.pipeline_stable_video_diffusion",
         tensor2vid=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("decode is outside the shim")),
         StableVideoDiffusionPipelineOutput=StableVideoDiffusionPipelineOutput, _append_dims=_append_dims,
         EXAMPLE_DOC_STRING="")
    _mod("diffusers.image_processor", VaeImageProcessor=VaeImageProcessor)
    del base
    # the reference package, without its __init__ files (they import training-only modules)
    _mod("ctrlv", _ctrlv_ref_shim=True)
    models = _mod("ctrlv.models")
    _mod("ctrlv.models.attention", BBOXFrameAttention=type("BBOXFrameAttention", (nn.Module,), {}))
    _mod("ctrlv.utils", get_fourier_embeds_from_boundingbox=None)
    u = _load("ctrlv.models.unet_spatio_temporal_condition", "models/unet_spatio_temporal_condition.py", "ctrlv.models")
    models.UNetSpatioTemporalConditionModel = u.UNetSpatioTemporalConditionModel
    c = _load("ctrlv.models.controlnet", "models/controlnet.py", "ctrlv.models")
    models.ControlNetModel = c.ControlNetModel
    _mod("ctrlv.pipelines")
    p = _load("ctrlv.pipelines.pipeline_video_control", "pipelines/pipeline_video_control.py", "ctrlv.pipelines")
    d = _load("ctrlv.pipelines.pipeline_video_diffusion", "pipelines/pipeline_video_diffusion.py", "ctrlv.pipelines")
    return SimpleNamespace(UNetSpatioTemporalConditionModel=u.UNetSpatioTemporalConditionModel,
                           ControlNetModel=c.ControlNetModel,
                           StableVideoControlPipeline=p.StableVideoControlPipeline,
                           VideoDiffusionPipeline=d.VideoDiffusionPipeline)


def unload():
    for k in [k for k in sys.modules if k.split(".")[0] in _SHIMMED
              and getattr(sys.modules[k.split(".")[0]], "_ctrlv_ref_shim", False)]:
        del sys.modules[k]
