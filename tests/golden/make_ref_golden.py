"""Generates tests/golden/ref_forward.pt by running the REFERENCE'S OWN forward / pipeline code
(loaded from /root/reference through `ref_shim.py`: diffusers blocks = the oracle's restatements).

Unlike `tiny_step.pt` (oracle pinned against itself), these vectors come out of the reference's
`ControlNetModel.forward`, `UNetSpatioTemporalConditionModel.forward` and
`StableVideoControlPipeline.__call__` / `VideoDiffusionPipeline.__call__` as they lie in the repo.
Only runs where /root/reference exists.  Re-run:  python tests/golden/make_ref_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import ref_shim as R  # noqa: E402
from oracle import svd_oracle as O  # noqa: E402

T, H, W, STEPS = 4, 8, 8, 25  # the reference default schedule (pipeline_video_control.py:112)
XDIM = O.TINY_CONFIG["cross_attention_dim"]


def oracle_models():
    torch.manual_seed(0)
    ou = O.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG).eval()
    oc = O.ControlNetModel(**O.TINY_CONFIG).eval()
    O.randomize_zero_convs(oc)
    return ou, oc


def reference_models(ref, ou, oc):
    """The reference's classes, constructed by the reference's constructors, with the oracle's weights."""
    ru = ref.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG).eval()
    rc = ref.ControlNetModel(**{k: v for k, v in O.TINY_CONFIG.items()}).eval()
    ru.load_state_dict(ou.state_dict(), strict=True)
    rc.load_state_dict(oc.state_dict(), strict=True)
    return ru, rc


def forward_inputs(batch=2, seed=11):
    g = torch.Generator().manual_seed(seed)
    return dict(sample=torch.randn(batch, T, 8, H, W, generator=g),
                timestep=torch.tensor(1.25),
                encoder_hidden_states=torch.randn(batch, 1, XDIM, generator=g),
                added_time_ids=torch.tensor([[6.0, 127.0, 0.02]]).repeat(batch, 1),
                control_cond=torch.randn(batch, T, 4, H, W, generator=g))


def pipeline_inputs(batch=1, seed=12, cond_channels=4):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(batch, 3, 8 * H, 8 * W, generator=g)
    cond = (torch.randn(batch, T, 4, H, W, generator=g) if cond_channels == 4
            else torch.rand(batch, T, 3, 8 * H, 8 * W, generator=g))
    latents = torch.randn(batch, T, 4, H, W, generator=g)
    return image, cond, latents


def run_control_pipeline(ref, ru, rc, image, cond, latents, scale=1.0, gen_seed=5, trace=None):
    pipe = ref.StableVideoControlPipeline(vae=R.FakeVAE(), image_encoder=R.FakeImageEncoder(XDIM), unet=ru,
                                          controlnet=rc, scheduler=R.SchedulerShim(), feature_extractor=None)
    cb = None if trace is None else (lambda p, i, t, kw: trace.append(kw["latents"].clone()) or {})
    return pipe(image, cond_images=cond, height=8 * H, width=8 * W, num_frames=T, num_inference_steps=STEPS,
                control_condition_scale=scale, generator=torch.Generator().manual_seed(gen_seed),
                latents=latents.clone(), output_type="latent", callback_on_step_end=cb).frames


def conditioning(image, cond, latents, gen_seed=5, noise_aug=0.02):
    """The conditioning `StableVideoControlPipeline.__call__` derives (pipeline_video_control.py:220-256,
    :71-101), computed independently of the reference code, in the `sampling.make_inputs` layout
    (+ the conditional halves alone, which is what this repo's pipeline takes as precomputed inputs)."""
    vae, enc = R.FakeVAE(), R.FakeImageEncoder(XDIM)
    B = image.shape[0]
    g = torch.Generator().manual_seed(gen_seed)
    noise = torch.randn(image.shape, generator=g, dtype=image.dtype)
    z = vae.latents(2.0 * image - 1.0 + noise_aug * noise)
    il = torch.cat([torch.zeros_like(z), z]).unsqueeze(1).repeat(1, T, 1, 1, 1)
    e = enc.embeds(image)
    if cond.shape[2] == 3:
        cond = vae.latents(cond.flatten(0, 1)).reshape(B, T, 4, H, W)
    return dict(latents=latents.clone(), image_latents=il, image_embeddings=torch.cat([torch.zeros_like(e), e]),
                cond_em=torch.cat([torch.zeros_like(cond), cond]),
                added_time_ids=torch.tensor([[6.0, 127.0, 0.02]]).repeat(2 * B, 1),
                guidance=torch.linspace(1.0, 3.0, T),
                image_latents_cond=z, image_embeds_cond=e, cond_em_cond=cond)


def run_bbox_pipeline(ref, ru, image, cond_frames, latents, gen_seed=5, num_cond_bbox_frames=1):
    pipe = ref.VideoDiffusionPipeline(vae=R.FakeVAE(), image_encoder=R.FakeImageEncoder(XDIM), unet=ru,
                                      scheduler=R.SchedulerShim(), feature_extractor=None)
    return pipe(image, bbox_images=cond_frames, height=8 * H, width=8 * W, num_frames=T,
                num_inference_steps=STEPS, generator=torch.Generator().manual_seed(gen_seed),
                latents=latents.clone(), output_type="latent", num_cond_bbox_frames=num_cond_bbox_frames).frames


def run():
    ref = R.load_reference()
    ou, oc = oracle_models()
    ru, rc = reference_models(ref, ou, oc)
    out = {"T": T, "H": H, "W": W, "steps": STEPS}
    with torch.no_grad():
        fi = forward_inputs()
        down, mid = rc(fi["sample"], fi["timestep"], fi["encoder_hidden_states"], fi["added_time_ids"],
                       control_cond=fi["control_cond"], conditioning_scale=0.7, return_dict=False)
        # the 12 residuals in full would be 660 KB: keep norm + sum of each and two of them whole
        out["ctrl_down_stats"] = torch.stack([torch.stack([d.norm(), d.sum()]) for d in down])
        out["ctrl_down_3"], out["ctrl_down_11"] = down[3].clone(), down[11].clone()
        out["ctrl_mid"] = mid.clone()
        out["unet_with_residuals"] = ru(fi["sample"], fi["timestep"], fi["encoder_hidden_states"],
                                        fi["added_time_ids"], down_block_additional_residuals=down,
                                        mid_block_additional_residuals=mid, return_dict=False)[0]
        out["unet_plain"] = ru(fi["sample"], fi["timestep"], fi["encoder_hidden_states"], fi["added_time_ids"]).sample
        image, cond, latents = pipeline_inputs()
        trace = []
        out["control_pipeline_latents"] = run_control_pipeline(ref, ru, rc, image, cond, latents, trace=trace)
        out["control_pipeline_trace"] = torch.stack(trace)  # latents after each of the STEPS steps
        image, condf, latents = pipeline_inputs(cond_channels=3, seed=13)
        out["control_pipeline_latents_from_frames"] = run_control_pipeline(ref, ru, rc, image, condf, latents, scale=0.5)
        out["bbox_pipeline_latents"] = run_bbox_pipeline(ref, ru, image, condf, latents)
    R.unload()
    return out


if __name__ == "__main__":
    o = run()
    path = os.path.join(HERE, "ref_forward.pt")
    torch.save(o, path)
    print("wrote", path, os.path.getsize(path), "bytes")
