"""ORACLE (test infrastructure — never imported by the product path).

Restates the sampling arithmetic of the Box2Video hot path:
  * EulerDiscreteScheduler of diffusers==0.27.2 with the SVD scheduler config
    (schedulers/scheduling_euler_discrete.py: set_timesteps / _convert_to_karras /
    scale_model_input / step) — called from
    /root/reference/src/ctrlv/pipelines/pipeline_video_control.py:259,301,332;
  * the loop body pipeline_video_control.py:298-343 (`denoise_step`) and the loop (`sample_loop`).
PARITY: the loops (`denoise_step`, `sample_loop`, `sample_loop_bbox_predictor`) are PINNED bit-exact against the
reference's own `__call__`s run through tests/golden/ref_shim.py (tests/test_reference_pin.py); the scheduler
arithmetic (a diffusers restatement) is UNPINNED — no reference fixtures exist — and held by the known answers
in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch


class EulerDiscreteSchedulerOracle:
    def __init__(self, sigma_min=0.002, sigma_max=700.0, rho=7.0, timestep_spacing="leading"):
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho
        self.timestep_spacing = timestep_spacing
        self.sigmas = None
        self.timesteps = None
        self.step_index = None

    def set_timesteps(self, n: int):
        # use_karras_sigmas with config sigma_min / sigma_max (numpy float64 -> float32)
        ramp = np.linspace(0, 1, n)
        min_inv_rho = self.sigma_min ** (1 / self.rho)
        max_inv_rho = self.sigma_max ** (1 / self.rho)
        sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** self.rho
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32)
        # timestep_type == "continuous" and prediction_type == "v_prediction"
        self.timesteps = torch.Tensor([0.25 * sigma.log() for sigma in sigmas])
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.step_index = None

    @property
    def init_noise_sigma(self):
        max_sigma = self.sigmas.max()
        if self.timestep_spacing in ("linspace", "trailing"):
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    def _init_step_index(self, t):
        idx = (self.timesteps == t).nonzero()
        self.step_index = idx[1 if len(idx) > 1 else 0].item()

    def scale_model_input(self, sample, t):
        if self.step_index is None:
            self._init_step_index(t)
        sigma = self.sigmas[self.step_index]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, t, sample):
        if self.step_index is None:
            self._init_step_index(t)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self.step_index]
        sigma_hat = sigma  # gamma = 0 (s_churn = 0)
        # v_prediction
        pred_original_sample = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        derivative = (sample - pred_original_sample) / sigma_hat
        dt = self.sigmas[self.step_index + 1] - sigma_hat
        prev_sample = sample + derivative * dt
        prev_sample = prev_sample.to(model_output.dtype)
        self.step_index += 1
        return prev_sample


def denoise_step(unet, controlnet, scheduler, latents, t, image_latents, image_embeddings, added_time_ids,
                 cond_em, guidance_scale, conditioning_scale=1.0, do_cfg=True, return_noise=False):
    """pipeline_video_control.py:298-332 for one timestep t."""
    latent_model_input = torch.cat([latents] * 2) if do_cfg else latents
    latent_model_input = scheduler.scale_model_input(latent_model_input, t)
    latent_model_input = torch.cat([latent_model_input, image_latents], dim=2)
    down = mid = None
    if controlnet is not None:
        down, mid = controlnet(latent_model_input, timestep=t, encoder_hidden_states=image_embeddings,
                               added_time_ids=added_time_ids, control_cond=cond_em,
                               conditioning_scale=conditioning_scale, return_dict=False)
    noise_pred_raw = unet(sample=latent_model_input, timestep=t, encoder_hidden_states=image_embeddings,
                          added_time_ids=added_time_ids, down_block_additional_residuals=down,
                          mid_block_additional_residuals=mid, return_dict=False)[0]
    noise_pred = noise_pred_raw
    if do_cfg:
        u, c = noise_pred.chunk(2)
        noise_pred = u + guidance_scale * (c - u)
    latents = scheduler.step(noise_pred, t, latents)
    if return_noise:
        return latents, noise_pred_raw
    return latents


def make_inputs(T=14, h=40, w=64, xdim=1024, seed=1234, batch=1, device="cpu"):
    """Synthetic inputs of SURVEY.md §8(d): drawn on the CPU in fp32 in a fixed order."""
    g = torch.Generator("cpu").manual_seed(seed)
    latents = torch.randn(batch, T, 4, h, w, generator=g)
    image_latents_cond = torch.randn(batch, 4, h, w, generator=g)
    image_embeds_cond = torch.randn(batch, 1, xdim, generator=g)
    cond_em_cond = torch.randn(batch, T, 4, h, w, generator=g)
    il = image_latents_cond.unsqueeze(1).repeat(1, T, 1, 1, 1)
    out = dict(
        latents=latents,
        image_latents=torch.cat([torch.zeros_like(il), il]),
        image_embeddings=torch.cat([torch.zeros_like(image_embeds_cond), image_embeds_cond]),
        cond_em=torch.cat([torch.zeros_like(cond_em_cond), cond_em_cond]),
        added_time_ids=torch.tensor([[6.0, 127.0, 0.02]]).repeat(2 * batch, 1),
        guidance=torch.linspace(1.0, 3.0, T),
        image_latents_cond=image_latents_cond, image_embeds_cond=image_embeds_cond, cond_em_cond=cond_em_cond,
    )
    return {k: v.to(device) for k, v in out.items()}


def sample_loop(unet, controlnet, inputs, num_steps=25, conditioning_scale=1.0, trace=None):
    """pipeline_video_control.py:259-343 with precomputed conditioning; returns final latents."""
    sch = EulerDiscreteSchedulerOracle()
    sch.set_timesteps(num_steps)
    dev = inputs["latents"].device
    latents = inputs["latents"] * sch.init_noise_sigma
    gs = inputs["guidance"].view(1, -1, 1, 1, 1).to(dev)
    for t in sch.timesteps:
        latents = denoise_step(unet, controlnet, sch, latents, t, inputs["image_latents"],
                               inputs["image_embeddings"], inputs["added_time_ids"], inputs["cond_em"], gs,
                               conditioning_scale)
        if trace is not None:
            trace.append(latents.clone())
    return latents


def sample_loop_bbox_predictor(unet, inputs, num_steps=25, num_cond_bbox_frames=3, trace=None):
    """/root/reference/src/ctrlv/pipelines/pipeline_video_diffusion.py:196-293 with precomputed
    conditioning: the plain SVD sampler of the bbox-predictor stage.  The repeated image latents are
    overwritten by the bbox-frame latents for the first `num_cond_bbox_frames` frames and the last
    frame (:200-206; the CFG-uncond half receives the zeros of `_encode_vae_condition`)."""
    sch = EulerDiscreteSchedulerOracle()
    sch.set_timesteps(num_steps)
    dev = inputs["latents"].device
    image_latents = inputs["image_latents"].clone()
    cond = inputs["cond_em"]
    image_latents[:, 0:num_cond_bbox_frames] = cond[:, 0:num_cond_bbox_frames]
    image_latents[:, -1] = cond[:, -1]
    latents = inputs["latents"] * sch.init_noise_sigma
    gs = inputs["guidance"].view(1, -1, 1, 1, 1).to(dev)
    for t in sch.timesteps:
        latents = denoise_step(unet, None, sch, latents, t, image_latents, inputs["image_embeddings"],
                               inputs["added_time_ids"], None, gs)
        if trace is not None:
            trace.append(latents.clone())
    return latents
