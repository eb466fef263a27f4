"""Sampling side of the hot path: the Euler-EDM schedule, the fused CFG denoise step (one CUDA
graph per shape) and a drop-in `StableVideoControlPipeline.__call__`.

Reference:
  /root/reference/src/ctrlv/pipelines/pipeline_video_control.py:105-360  (__call__)
  loop body :298-343 — cat([latents]*2), scale_model_input, cat(image_latents), ControlNet, UNet,
  CFG combine, scheduler.step.
Scheduler arithmetic: diffusers==0.27.2 EulerDiscreteScheduler with the SVD config.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import torch

from . import ops
from .models import BF16, ControlNetModel, UNetSpatioTemporalConditionModel


class EulerDiscreteScheduler:
    """Karras-sigma, v-prediction, continuous-timestep Euler scheduler (SVD configuration).

    Host-side schedule construction only (numpy); the per-step arithmetic
    (`scale_model_input`, `step`) runs in the fused CUDA kernels of this package."""

    order = 1

    def __init__(self, sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0,
                 num_train_timesteps: int = 1000, timestep_spacing: str = "leading"):
        self.config = SimpleNamespace(sigma_min=sigma_min, sigma_max=sigma_max, rho=rho,
                                      num_train_timesteps=num_train_timesteps,
                                      prediction_type="v_prediction", timestep_type="continuous",
                                      use_karras_sigmas=True, timestep_spacing=timestep_spacing,
                                      steps_offset=1)
        self.sigmas = None
        self.timesteps = None
        self._step_index = None
        self.num_inference_steps = None

    @classmethod
    def from_config(cls, config: dict) -> "EulerDiscreteScheduler":
        """From a diffusers `scheduler_config.json` dictionary.  Only the mode the fused CFG + Euler
        kernel implements is accepted (the SVD configuration: v-prediction, continuous timesteps,
        Karras sigmas); anything else raises instead of silently sampling with other arithmetic."""
        want = {"prediction_type": "v_prediction", "timestep_type": "continuous", "use_karras_sigmas": True}
        for k, v in want.items():
            if config.get(k, v) != v:
                raise NotImplementedError(f"scheduler config {k}={config[k]!r}: only {k}={v!r} (the SVD "
                                          "EulerDiscreteScheduler configuration) runs on the sm_100a path")
        name = config.get("_class_name", "EulerDiscreteScheduler")
        if name != "EulerDiscreteScheduler":
            raise NotImplementedError(f"scheduler class {name!r}: only EulerDiscreteScheduler is implemented")
        return cls(sigma_min=float(config.get("sigma_min", 0.002)), sigma_max=float(config.get("sigma_max", 700.0)),
                   num_train_timesteps=int(config.get("num_train_timesteps", 1000)),
                   timestep_spacing=config.get("timestep_spacing", "leading"))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, **_unused):
        """`SchedulerMixin.from_pretrained` for a local directory holding `scheduler_config.json`."""
        import json
        import os
        d = os.path.join(pretrained_model_name_or_path, subfolder) if subfolder else pretrained_model_name_or_path
        path = os.path.join(d, "scheduler_config.json")
        if not os.path.isfile(path):
            raise OSError(f"{path} not found (this build does not download from the hub)")
        with open(path) as f:
            return cls.from_config(json.load(f))

    def set_timesteps(self, num_inference_steps: int, device=None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho = c.sigma_min ** (1 / c.rho)
        max_inv_rho = c.sigma_max ** (1 / c.rho)
        sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** c.rho  # float64
        sig32 = torch.from_numpy(sigmas).to(torch.float32)
        self.timesteps = torch.tensor([0.25 * math.log(float(s)) for s in sig32], dtype=torch.float32)
        self.sigmas = torch.cat([sig32, torch.zeros(1, dtype=torch.float32)])  # kept on the host
        self._step_index = None
        return self

    @property
    def init_noise_sigma(self) -> float:
        m = float(self.sigmas.max())
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return m
        return (m ** 2 + 1) ** 0.5

    def index_for_timestep(self, t) -> int:
        t = float(t)
        d = (self.timesteps - t).abs()
        return int(torch.argmin(d).item())


class DenoiseStep:
    """One CFG Euler-EDM iteration of ControlNet + UNet on a fixed shape, replayed from a CUDA
    graph.  State: `latents` fp32 [B, T, 4, h, w] updated in place on the device.

    `cfg_branch` = 0 / 1 is the CFG-branch-sharded mode: this process computes only
    the uncond / cond half of the CFG batch; `exchange(step)` must then fill the other half of
    `step.noise` (an NCCL all-gather inside the GPU pair, `parallel.CfgPair`) before the CFG
    combine + Euler update, which both processes of a pair run redundantly on identical inputs."""

    def __init__(self, unet: UNetSpatioTemporalConditionModel, controlnet: Optional[ControlNetModel],
                 batch: int, num_frames: int, h: int, w: int, cfg: bool = True,
                 conditioning_scale: float = 1.0, use_graph: bool = True, two_streams: bool = True,
                 cfg_branch: Optional[int] = None, exchange: Optional[Callable] = None, use_plan: bool = False):
        if cfg_branch is not None and (not cfg or cfg_branch not in (0, 1)):
            raise ValueError("cfg_branch is 0 (uncond) or 1 (cond) and needs cfg=True")
        self.cfg_branch, self.exchange = cfg_branch, exchange
        self.unet, self.controlnet = unet, controlnet
        self.two_streams = two_streams
        self._side = torch.cuda.Stream() if two_streams else None
        self.B, self.T, self.h, self.w, self.cfg = batch, num_frames, h, w, cfg
        self.nb = 2 * batch if cfg else batch
        self.scale = conditioning_scale
        self.use_graph = use_graph
        # use_plan: replay the step from a library launch plan (ctrlv_plan_run: the launches are re-issued from C
        # with the arguments recorded once) instead of a CUDA graph — what a non-Python host would do
        self.use_plan = use_plan
        self._plan, self._pool = None, None
        dev = "cuda"
        xdim = unet.cfg["cross_attention_dim"]
        xdim = xdim if isinstance(xdim, int) else xdim[0]
        self.latents = torch.zeros((batch, num_frames, 4, h, w), device=dev, dtype=torch.float32)
        self.image_latents = torch.zeros((self.nb, num_frames, 4, h, w), device=dev, dtype=torch.float32)
        self.cond_em = torch.zeros((self.nb, num_frames, 4, h, w), device=dev, dtype=torch.float32)
        self.ehs = torch.zeros((self.nb, xdim), device=dev, dtype=torch.float32)
        self.added_time_ids = torch.zeros((self.nb, 3), device=dev, dtype=torch.float32)
        self.guidance = torch.ones((num_frames,), device=dev, dtype=torch.float32)
        # per-step scalars [sigma, sigma_next, t x nb], refreshed by one small H2D copy per step
        self.step_params = torch.zeros((2 + self.nb,), device=dev, dtype=torch.float32)
        self.noise = torch.zeros((self.nb * num_frames * h * w, unet.cfg["out_channels"]), device=dev,
                                 dtype=torch.float32)
        nloc = batch if cfg_branch is not None else self.nb  # samples this process pushes through the nets
        self.inp = torch.zeros((nloc * num_frames * h * w, 64), device=dev, dtype=BF16)
        # branch-sharded: the local half of the model output, sent to the peer by `exchange`
        self.noise_local = (torch.zeros((nloc * num_frames * h * w, unet.cfg["out_channels"]), device=dev,
                                        dtype=torch.float32) if cfg_branch is not None else None)
        self._graph = None
        self._host_params = None
        self.launches_per_step = None

    # -- the work of one step (kernel launches only) ------------------------------------------------
    def _body(self):
        self._body_model()
        if self.cfg_branch is None:
            self._body_update()

    def _body_update(self):
        ops.cfg_euler(self.latents, self.noise, self.cfg, self.guidance, self.step_params[0:2])

    def _body_model(self):
        sig = self.step_params[0:2]
        br = self.cfg_branch
        if br is None:
            g = (self.nb, self.T, self.h, self.w)
            ts, ids, out = self.step_params[2:], self.added_time_ids, self.noise
            il, ce, dup = self.image_latents, self.cond_em, self.cfg
        else:  # this process owns global batch rows br*B .. br*B + B - 1
            lo, hi = br * self.B, (br + 1) * self.B
            g = (self.B, self.T, self.h, self.w)
            ts, ids, out = self.step_params[2 + lo:2 + hi], self.added_time_ids[lo:hi], self.noise_local
            il, ce, dup = self.image_latents[lo:hi], self.cond_em[lo:hi], False
        ops.prep_input(self.latents, il, ce if self.controlnet is not None else None, dup, sig, out=self.inp)
        down = mid = None
        join = None
        if self.controlnet is not None:
            if self.two_streams:
                # the ControlNet and the UNet encoder both depend only on `inp` (they meet at
                # unet_spatio_temporal_condition.py:119): run the ControlNet on a side stream so its
                # kernels fill the tails / small grids of the UNet encoder's
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                ops.lib().ctrlv_plan_fork()  # (recorded when a launch plan is being built, else a no-op)
                with torch.cuda.stream(self._side):
                    emb_c = self.controlnet.embed(ts, ids)
                    down, mid, _, _ = self.controlnet.forward_rows(self.inp, emb_c, self.ehs, g, self.scale, branch=br)
                    for t in list(down) + [mid]:
                        t.record_stream(main)
                join = lambda: (main.wait_stream(self._side), ops.lib().ctrlv_plan_join())
            else:
                emb_c = self.controlnet.embed(ts, ids)
                down, mid, _, _ = self.controlnet.forward_rows(self.inp, emb_c, self.ehs, g, self.scale, branch=br)
        emb_u = self.unet.embed(ts, ids)
        self.unet.forward_rows(self.inp, emb_u, self.ehs, g, down, mid, out_f32=out, join=join, branch=br)

    def __del__(self):
        if getattr(self, "_plan", None) is not None:
            try:
                ops.lib().ctrlv_plan_destroy(self._plan)
            except Exception:
                pass
            self._plan = None

    def set_schedule(self, sigmas: torch.Tensor, timesteps: torch.Tensor):
        n = timesteps.numel()
        hp = torch.empty((n, 2 + self.nb), dtype=torch.float32).pin_memory()
        hp[:, 0] = sigmas[:n]
        hp[:, 1] = sigmas[1:n + 1]
        hp[:, 2:] = timesteps[:, None]
        self._host_params = hp

    def capture(self):
        """Warm up eagerly (first-use attribute setup, embedding caches), then capture."""
        if self._host_params is None:
            raise RuntimeError("set_schedule() first")
        keep = self.latents.clone()
        self.step_params.copy_(self._host_params[0], non_blocking=True)
        self._body()
        torch.cuda.synchronize()
        if self.use_plan:
            # record one step into a library plan; its intermediates come from a private pool that is never used
            # again, so the recorded addresses stay this step's own
            import ctypes
            from ._lib import check
            lib = ops.lib()
            self._pool = torch.cuda.MemPool()
            plan = ctypes.c_void_p()
            with torch.cuda.use_mem_pool(self._pool):
                check(lib.ctrlv_plan_create(torch.cuda.current_stream().cuda_stream, ctypes.byref(plan)), "ctrlv_plan_create")
                try:
                    self._body()
                finally:
                    check(lib.ctrlv_plan_finish(plan), "ctrlv_plan_finish")
            self._plan = plan
            torch.cuda.synchronize()
        elif self.use_graph:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                self._body()
            self._graph = gr
        self.latents.copy_(keep)
        torch.cuda.synchronize()

    def run_model(self, i: int):
        """Everything of step i up to the model output (the whole step when not branch-sharded)."""
        self.step_params.copy_(self._host_params[i], non_blocking=True)
        if self._plan is not None:
            from ._lib import check
            check(ops.lib().ctrlv_plan_run(self._plan, torch.cuda.current_stream().cuda_stream), "ctrlv_plan_run")
        elif self._graph is not None:
            self._graph.replay()
        else:
            self._body()

    def finish(self):
        """Branch-sharded mode: fetch the peer's half of the model output, then CFG + Euler."""
        if self.cfg_branch is None:
            return
        M = self.noise_local.shape[0]
        self.noise[self.cfg_branch * M:(self.cfg_branch + 1) * M].copy_(self.noise_local)
        if self.exchange is None:
            raise RuntimeError("branch-sharded DenoiseStep needs an `exchange` callable (parallel.CfgPair.exchange)")
        self.exchange(self)
        self._body_update()

    def step(self, i: int):
        self.run_model(i)
        self.finish()

    # -- host-buffer entry point (what a reference-side caller holding host tensors would use) -----
    HOST_INPUTS = ("latents", "image_latents", "cond_em", "ehs", "added_time_ids", "guidance")

    def make_host_buffers(self) -> Dict[str, torch.Tensor]:
        """Pinned host mirrors of every per-step input plus the output latents."""
        hb = {k: torch.empty(getattr(self, k).shape, dtype=torch.float32).pin_memory() for k in self.HOST_INPUTS}
        hb["latents_out"] = torch.empty(self.latents.shape, dtype=torch.float32).pin_memory()
        return hb

    def step_host(self, i: int, hb: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One denoise step with HOST inputs/outputs: H2D of the step's inputs, the step, D2H of the
        updated latents, stream sync.  Returns hb["latents_out"]."""
        for k in self.HOST_INPUTS:
            getattr(self, k).copy_(hb[k], non_blocking=True)
        self.step(i)
        hb["latents_out"].copy_(self.latents, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return hb["latents_out"]

    def host_bytes_per_step(self):
        h2d = sum(getattr(self, k).numel() * 4 for k in self.HOST_INPUTS) + self.step_params.numel() * 4
        return h2d, self.latents.numel() * 4


class StableVideoDiffusionPipelineOutput(SimpleNamespace):
    pass


class StableVideoControlPipeline:
    """Drop-in for ctrlv.pipelines.StableVideoControlPipeline (pipeline_video_control.py:25-360).

    The denoising loop (:298-343) runs on the sm_100a kernels.  With a `vae`
    (`ctrlv_b200.vae.AutoencoderKLTemporalDecoder`) the pipeline also encodes
    3-channel bbox frames (:84) and the conditioning image (:235) and decodes the result (:346-347,
    `output_type` "pt" / "np" / "pil"); with an `image_encoder`
    (`ctrlv_b200.clip.CLIPVisionModelWithProjection`, row f-3) it embeds the conditioning image (:220:
    antialiased resize to 224 + CLIP ViT-H).  `image` is a PIL image, a list of them, or a
    [B, 3, H, W] tensor in [0, 1] (resized to `height` x `width` like `VaeImageProcessor.preprocess`).  Without those modules pass
    `image_embeddings=`, `image_latents=` and 4-channel `cond_images` latents (:86-88) and use
    `output_type="latent"`."""

    def __init__(self, vae=None, image_encoder=None, unet: UNetSpatioTemporalConditionModel = None,
                 controlnet: ControlNetModel = None, scheduler: EulerDiscreteScheduler = None,
                 feature_extractor=None):
        if unet is None:
            raise ValueError("`unet` is required")
        self.vae, self.image_encoder, self.feature_extractor = vae, image_encoder, feature_extractor
        self.unet, self.controlnet = unet, controlnet
        self.scheduler = scheduler if scheduler is not None else EulerDiscreteScheduler()
        self.vae_scale_factor = (2 ** (len(vae.config.block_out_channels) - 1)) if vae is not None else 8
        self._steps: Dict[tuple, DenoiseStep] = {}  # insertion-ordered: least recently used first
        self.max_cached_steps = 4
        self._guidance_scale = 1.0

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, **kwargs):
        """`DiffusionPipeline.from_pretrained` for a LOCAL Stable-Video-Diffusion directory
        (tools/eval_video_controlnet.py:116-118): loads `vae/` and `image_encoder/` (and `unet/` unless a
        `unet=` is passed) with the sm_100a drop-ins; `controlnet=` / `unet=` keyword modules are used as
        given.  The scheduler comes from `scheduler/scheduler_config.json` when the directory has one (it must
        be the SVD EulerDiscreteScheduler mode), else it is the SVD default configuration.  No hub download."""
        import os
        from . import clip as _clip
        from . import vae as _vae
        root = pretrained_model_name_or_path
        if not os.path.isdir(root):
            raise OSError(f"{root} is not a local directory (this build does not download from the hub)")
        kwargs.pop("torch_dtype", None); variant = kwargs.pop("variant", None)
        parts = {}
        for name, loader in (("vae", _vae.AutoencoderKLTemporalDecoder), ("image_encoder", _clip.CLIPVisionModelWithProjection),
                             ("unet", UNetSpatioTemporalConditionModel), ("controlnet", ControlNetModel)):
            if name in kwargs:
                parts[name] = kwargs.pop(name)
            elif os.path.isdir(os.path.join(root, name)):
                parts[name] = loader.from_pretrained(root, subfolder=name, variant=variant)
            else:
                parts[name] = None
        if cls is VideoDiffusionPipeline:
            parts.pop("controlnet", None)
        scheduler = kwargs.pop("scheduler", None)
        if scheduler is None and os.path.isfile(os.path.join(root, "scheduler", "scheduler_config.json")):
            scheduler = EulerDiscreteScheduler.from_pretrained(root, subfolder="scheduler")
        return cls(scheduler=scheduler, feature_extractor=kwargs.pop("feature_extractor", None), **parts)

    def to(self, *a, **k):  # modules are device-resident by construction
        return self

    def set_progress_bar_config(self, **k):
        pass

    @property
    def do_classifier_free_guidance(self):
        g = self._guidance_scale
        return (g > 1) if isinstance(g, (int, float)) else bool(g.max() > 1)

    def check_inputs(self, image, cond_images, height, width):  # pipeline_video_control.py:51-68
        # `image=None` is this package's extension (precomputed image_embeddings= / image_latents=)
        if image is not None and not isinstance(image, (torch.Tensor, list)):
            import PIL.Image
            if not isinstance(image, PIL.Image.Image):
                raise ValueError("`image` has to be of type `torch.FloatTensor` or `PIL.Image.Image` or "
                                 f"`List[PIL.Image.Image]` but is {type(image)}")
        if not isinstance(cond_images, torch.Tensor):
            raise ValueError("`cond_images` has to be of type `torch.FloatTensor` but is " f"{type(cond_images)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _encode_vae_condition(self, cond_image, num_videos_per_prompt, do_cfg):  # :71-101
        if cond_image.shape[2] == 3:  # frames -> latents through the VAE encoder, `.mode()` (:84)
            if self.vae is None:
                raise NotImplementedError("3-channel cond_images need a `vae`; pass 4-channel bbox-frame latents")
            b, f = cond_image.shape[:2]
            em = self.vae.encode(cond_image.to("cuda", torch.float32).flatten(0, 1)).latent_dist.mode()
            cond_image = em.reshape(b, f, *em.shape[1:])
        assert cond_image.shape[2] == 4, "The input tensor should have 3 or 4 channels. 3 for frames and 4 for latents."
        cond_em = cond_image.to("cuda", torch.float32).repeat(num_videos_per_prompt, 1, 1, 1, 1)
        if do_cfg:
            cond_em = torch.cat([torch.zeros_like(cond_em), cond_em])
        return cond_em

    @staticmethod
    def _image_to_tensor01(image) -> torch.Tensor:
        """PIL image / list of PIL images / [B, 3, H, W] tensor -> float tensor in [0, 1]
        (`VaeImageProcessor.pil_to_numpy` + `numpy_to_pt`)."""
        if isinstance(image, torch.Tensor):
            if image.ndim != 4 or image.shape[1] != 3:
                raise ValueError(f"`image` tensor must be [B, 3, H, W] in [0, 1], got {tuple(image.shape)}")
            return image.to(torch.float32)
        import numpy as np
        imgs = image if isinstance(image, list) else [image]
        arr = np.stack([np.array(im.convert("RGB")).astype("float32") / 255.0 for im in imgs])
        return torch.from_numpy(arr).permute(0, 3, 1, 2).contiguous()

    def _encode_image(self, image) -> torch.Tensor:  # diffusers `_encode_image`, called at :220
        """-> conditional image embeddings [B, 1, D] (the CFG zeros are prepended by the caller)."""
        if self.image_encoder is None:
            raise NotImplementedError("pass image_embeddings=[B,1,D] or construct the pipeline with an "
                                      "`image_encoder` (ctrlv_b200.clip.CLIPVisionModelWithProjection)")
        return self.image_encoder.encode_image(self._image_to_tensor01(image))

    @staticmethod
    def _duplicate_conditioning(image_embeddings, image_latents, batch_size, nvp):
        """Per-prompt duplication for `num_videos_per_prompt` > 1, in the orders the reference's helpers
        use: the bbox-frame latents (`_encode_vae_condition`, pipeline_video_control.py:90) and the image
        latents (diffusers-0.27.2 `_encode_vae_image`: `.repeat(num_videos_per_prompt, 1, 1, 1)`) are
        TILED, so sample j carries image j % batch and bbox frames j % batch; only the CLIP embedding
        (`_encode_image`: `repeat(1, n, 1).view(b * n, ...)`) is interleaved.  For batch > 1 and n > 1 the
        reference therefore pairs embedding j // n with latents j % batch; that quirk is kept so that the
        drop-in samples what the reference samples."""
        emb = image_embeddings.to("cuda", torch.float32).reshape(batch_size, -1).repeat_interleave(nvp, 0)
        il = image_latents.to("cuda", torch.float32).repeat(nvp, 1, 1, 1)
        return emb, il

    def _denoise_step(self, key, make):
        """Cached `DenoiseStep` per (shape, mode, weight versions).  A captured graph holds raw pointers
        of the packed weights, so the versions of both networks are part of the key (a `load_state_dict`
        repacks and frees the old buffers); the cache is a small LRU because every entry owns a full
        activation pool."""
        key = key + (self.unet._version, self.controlnet._version if self.controlnet is not None else -1)
        st = self._steps.pop(key, None)
        if st is None:
            st = make()
            while len(self._steps) >= self.max_cached_steps:
                self._steps.pop(next(iter(self._steps)))
        self._steps[key] = st  # most recently used last
        return st

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, batch_size):  # :247-255
        add_time_ids = [fps, motion_bucket_id, noise_aug_strength]
        passed = self.unet.config.addition_time_embed_dim * len(add_time_ids)
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(
                f"Model expects an added time embedding vector of length {expected}, but a vector of "
                f"{passed} was created. The model has an incorrect config. Please check "
                "`unet.config.time_embedding_type` and `text_encoder_2.config.projection_dim`.")
        return torch.tensor([add_time_ids], dtype=torch.float32).repeat(batch_size, 1)

    @classmethod
    def _preprocess_image(cls, image, height: int, width: int) -> torch.Tensor:
        """`VaeImageProcessor.preprocess(image, height, width)` of diffusers 0.27.2 up to (not including) the
        [-1, 1] normalisation (pipeline_video_control.py:227): PIL images are resized with PIL's Lanczos filter
        (the processor's default `resample`), tensors with `F.interpolate(size=...)`'s default nearest-neighbour
        rule (source index = floor(dst * in / out)) — a pure gather, done with index_select.  Returns
        [B, 3, height, width] fp32 in [0, 1]."""
        if not isinstance(image, torch.Tensor):
            from PIL import Image
            imgs = image if isinstance(image, list) else [image]
            imgs = [im if im.size == (width, height) else im.resize((width, height), resample=Image.LANCZOS) for im in imgs]
            return cls._image_to_tensor01(imgs)
        image = cls._image_to_tensor01(image)
        H, W = image.shape[-2:]
        if (H, W) != (height, width):
            iy = torch.floor(torch.arange(height, dtype=torch.float32) * (H / height)).long().clamp_(max=H - 1)
            ix = torch.floor(torch.arange(width, dtype=torch.float32) * (W / width)).long().clamp_(max=W - 1)
            image = image.index_select(-2, iy.to(image.device)).index_select(-1, ix.to(image.device))
        return image

    def _encode_vae_image(self, image, height, width, noise_aug_strength, generator):  # :227-241
        """`image_processor.preprocess` (resize to height x width, normalise to [-1, 1]), noise augmentation,
        VAE `.mode()`."""
        if self.vae is None:
            raise NotImplementedError("pass image_latents=[B,4,h,w] or construct the pipeline with a `vae`")
        image = self._preprocess_image(image, height, width)
        image = 2.0 * image.to(torch.float32) - 1.0
        noise = torch.randn(image.shape, generator=generator, dtype=torch.float32,
                            device=generator.device if generator is not None else "cpu")
        image = image.to("cuda") + noise_aug_strength * noise.to("cuda")
        return self.vae.encode(image).latent_dist.mode()

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):  # diffusers decode_latents (:346)
        if self.vae is None:
            raise NotImplementedError("decoding needs a `vae`; use output_type='latent'")
        from . import vae as _vae
        return _vae.decode_latents(self.vae, latents, num_frames, decode_chunk_size)

    @staticmethod
    def tensor2vid(video: torch.Tensor, output_type: str = "np"):
        """diffusers `tensor2vid` + `VaeImageProcessor.postprocess`: [B, C, T, H, W] in [-1, 1] ->
        "pt": [B, T, C, H, W] in [0, 1]; "np": [B, T, H, W, C]; "pil": list of lists of PIL images."""
        if output_type not in ("pt", "np", "pil"):
            raise ValueError(f"output_type {output_type!r} is not one of 'latent', 'pt', 'np', 'pil'")
        vid = (video.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return vid
        arr = vid.permute(0, 1, 3, 4, 2).float().cpu().numpy()
        if output_type == "np":
            return arr
        from PIL import Image
        return [[Image.fromarray((fr * 255).round().astype("uint8")) for fr in clip] for clip in arr]

    def _finish(self, latents, num_frames, decode_chunk_size, output_type, return_dict, clamp=False):
        if output_type == "latent":
            frames = latents
        else:
            frames = self.decode_latents(latents, num_frames, decode_chunk_size or num_frames)
            if clamp:  # pipeline_video_diffusion.py:297
                frames = torch.clamp(frames, -1, 1)
            frames = self.tensor2vid(frames, output_type)
        if not return_dict:
            return frames
        return StableVideoDiffusionPipelineOutput(frames=frames)

    @torch.no_grad()
    def __call__(self, image=None, cond_images: torch.Tensor = None, height: int = 576, width: int = 1024,
                 num_frames: Optional[int] = None, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                 control_condition_scale: float = 1.0, fps: int = 7, motion_bucket_id: int = 127,
                 noise_aug_strength: float = 0.02, decode_chunk_size: Optional[int] = None,
                 num_videos_per_prompt: Optional[int] = 1, generator=None,
                 latents: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 image_embeddings: Optional[torch.Tensor] = None,
                 image_latents: Optional[torch.Tensor] = None, use_graph: bool = True):
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        self.check_inputs(image, cond_images, height, width)
        if image_embeddings is None:
            image_embeddings = self._encode_image(image)
        if output_type != "latent" and self.vae is None:
            raise NotImplementedError("decoding needs a `vae` (ctrlv_b200.vae.AutoencoderKLTemporalDecoder); "
                                      "use output_type='latent'")
        if image_latents is None:
            image_latents = self._encode_vae_image(image, height, width, noise_aug_strength, generator)
        batch_size = image_embeddings.shape[0]
        nvp = num_videos_per_prompt
        self._guidance_scale = max_guidance_scale  # :217 — CFG on iff > 1
        do_cfg = self.do_classifier_free_guidance
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        B = batch_size * nvp
        # :220 uncond image embedding = zeros; :235-245 uncond image latents = zeros
        emb, il = self._duplicate_conditioning(image_embeddings, image_latents, batch_size, nvp)
        if do_cfg:
            emb = torch.cat([torch.zeros_like(emb), emb])
            il = torch.cat([torch.zeros_like(il), il])
        il = il.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)
        fps = fps - 1  # :224
        ids = self._get_add_time_ids(fps, motion_bucket_id, noise_aug_strength, B)
        if do_cfg:
            ids = torch.cat([ids, ids])
        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = self.scheduler.timesteps
        shape = (B, num_frames, 4, h, w)
        if latents is None:
            latents = torch.randn(shape, generator=generator, dtype=torch.float32,
                                  device=generator.device if generator is not None else "cpu")
        latents = latents.to("cuda", torch.float32) * self.scheduler.init_noise_sigma
        cond_em = self._encode_vae_condition(cond_images, nvp, do_cfg)
        guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames)  # :287

        def make():
            st = DenoiseStep(self.unet, self.controlnet, B, num_frames, h, w, do_cfg,
                             control_condition_scale, use_graph)
            st.set_schedule(self.scheduler.sigmas, timesteps)
            st.capture()
            return st
        st = self._denoise_step((B, num_frames, h, w, do_cfg, float(control_condition_scale), num_inference_steps,
                                 use_graph), make)
        st.latents.copy_(latents)
        st.image_latents.copy_(il)
        st.cond_em.copy_(cond_em)
        st.ehs.copy_(emb)
        st.added_time_ids.copy_(ids.to("cuda"))
        st.guidance.copy_(guidance.to("cuda"))
        for i, t in enumerate(timesteps):
            st.step(i)
            if callback_on_step_end is not None:
                kw = {k: st.latents for k in callback_on_step_end_tensor_inputs if k == "latents"}
                out = callback_on_step_end(self, i, t, kw)
                if out and "latents" in out:
                    st.latents.copy_(out["latents"])
        return self._finish(st.latents.clone(), num_frames, decode_chunk_size, output_type, return_dict)


class VideoDiffusionPipeline(StableVideoControlPipeline):
    """Drop-in for ctrlv.pipelines.VideoDiffusionPipeline (pipeline_video_diffusion.py:19-330): the
    plain SVD sampler of the bbox-predictor stage.  Same loop as the control
    pipeline without the ControlNet (:259-293); the conditioning-frame overwrite (:200-206) puts the
    bbox-frame latents of the first `num_cond_bbox_frames` frames and of the last frame in place of
    the repeated image latents.  `bbox_images` are frames [B, T, 3, H, W] (encoded by the `vae`) or
    4-channel latents [B, T, 4, h, w]; decoded frames are clamped to [-1, 1] (:297)."""

    def __init__(self, vae=None, image_encoder=None, unet: UNetSpatioTemporalConditionModel = None,
                 scheduler: EulerDiscreteScheduler = None, feature_extractor=None):
        super().__init__(vae=vae, image_encoder=image_encoder, unet=unet, controlnet=None,
                         scheduler=scheduler, feature_extractor=feature_extractor)

    def check_inputs(self, image, height, width):  # diffusers StableVideoDiffusionPipeline.check_inputs
        if image is not None and not isinstance(image, (torch.Tensor, list)):  # None: precomputed conditioning
            import PIL.Image
            if not isinstance(image, PIL.Image.Image):
                raise ValueError("`image` has to be of type `torch.FloatTensor` or `PIL.Image.Image` or "
                                 f"`List[PIL.Image.Image]` but is {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    @torch.no_grad()
    def __call__(self, image=None, bbox_images: Optional[torch.Tensor] = None, bbox_conditions=None,
                 original_size=(1242, 375), height: int = 576, width: int = 1024,
                 num_frames: Optional[int] = None, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0, fps: int = 7,
                 motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                 decode_chunk_size: Optional[int] = None, num_videos_per_prompt: Optional[int] = 1,
                 generator=None, latents: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 num_cond_bbox_frames: int = 3, image_embeddings: Optional[torch.Tensor] = None,
                 image_latents: Optional[torch.Tensor] = None, use_graph: bool = True):
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        self.check_inputs(image, height, width)
        if image_embeddings is None:
            image_embeddings = self._encode_image(image)
        if output_type != "latent" and self.vae is None:
            raise NotImplementedError("decoding needs a `vae` (ctrlv_b200.vae.AutoencoderKLTemporalDecoder); "
                                      "use output_type='latent'")
        if image_latents is None:
            image_latents = self._encode_vae_image(image, height, width, noise_aug_strength, generator)
        batch_size = image_embeddings.shape[0]
        nvp = num_videos_per_prompt
        self._guidance_scale = max_guidance_scale
        do_cfg = self.do_classifier_free_guidance
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        B = batch_size * nvp
        emb, il = self._duplicate_conditioning(image_embeddings, image_latents, batch_size, nvp)
        if do_cfg:
            emb = torch.cat([torch.zeros_like(emb), emb])
            il = torch.cat([torch.zeros_like(il), il])
        il = il.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)  # :196
        if bbox_images is not None:  # :199-206
            cond = self._encode_vae_condition(bbox_images, nvp, do_cfg)
            il[:, 0:num_cond_bbox_frames] = cond[:, 0:num_cond_bbox_frames]
            il[:, -1] = cond[:, -1]
        ids = self._get_add_time_ids(fps - 1, motion_bucket_id, noise_aug_strength, B)
        if do_cfg:
            ids = torch.cat([ids, ids])
        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = self.scheduler.timesteps
        if latents is None:
            latents = torch.randn((B, num_frames, 4, h, w), generator=generator, dtype=torch.float32,
                                  device=generator.device if generator is not None else "cpu")
        latents = latents.to("cuda", torch.float32) * self.scheduler.init_noise_sigma
        guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames)  # :249

        def make():
            st = DenoiseStep(self.unet, None, B, num_frames, h, w, do_cfg, 1.0, use_graph)
            st.set_schedule(self.scheduler.sigmas, timesteps)
            st.capture()
            return st
        st = self._denoise_step((B, num_frames, h, w, do_cfg, num_inference_steps, use_graph), make)
        st.latents.copy_(latents)
        st.image_latents.copy_(il)
        st.ehs.copy_(emb)
        st.added_time_ids.copy_(ids.to("cuda"))
        st.guidance.copy_(guidance.to("cuda"))
        for i, t in enumerate(timesteps):
            st.step(i)
            if callback_on_step_end is not None:
                kw = {k: st.latents for k in callback_on_step_end_tensor_inputs if k == "latents"}
                out = callback_on_step_end(self, i, t, kw)
                if out and "latents" in out:
                    st.latents.copy_(out["latents"])
        return self._finish(st.latents.clone(), num_frames, decode_chunk_size, output_type, return_dict, clamp=True)
