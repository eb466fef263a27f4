"""GPU parity of the image-conditioning prologue: antialiased resize against the
reference's golden vectors, CLIP vision tower against the oracle (itself pinned to `transformers`)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def test_resize_with_antialiasing_matches_reference_golden():
    from ctrlv_b200 import clip
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "resize_antialias.pt"))
    for c in cases:
        got = clip.resize_with_antialiasing(c["input"].to(dev).contiguous(), c["size"])
        assert got.shape == c["output"].shape
        err = float((got.cpu() - c["output"]).abs().max())
        assert err < 5e-6, err


def test_clip_patchify_matches_unfold():
    from ctrlv_b200 import ops
    g = torch.Generator("cpu").manual_seed(0)
    img = torch.rand(2, 3, 28, 42, generator=g).to(dev)
    mean = torch.tensor([0.4, 0.5, 0.6], device=dev); std = torch.tensor([0.2, 0.3, 0.25], device=dev)
    rows = ops.clip_patchify(img, 14, mean, std, a=0.5, s=0.25, clamp01=True)
    x = ((0.5 * img + 0.25).clamp(0, 1) - mean.view(1, 3, 1, 1)) / std.view(1, 3, 1, 1)
    want = torch.nn.functional.unfold(x, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert rows.shape == (2 * 2 * 3, 640)
    assert float((rows[:, :588].float() - want).abs().max()) < 2e-2 and float(rows[:, 588:].abs().max()) == 0


def _pair(cfg, seed=0):
    from ctrlv_b200 import clip
    from oracle import clip_oracle as CO
    torch.manual_seed(seed)
    oc = CO.CLIPVisionModelWithProjection(**cfg).to(dev).eval()
    mc = clip.CLIPVisionModelWithProjection(state_dict=oc.state_dict(), **cfg)
    return oc, mc


def test_tiny_clip_tower_and_encode_image():
    from oracle import clip_oracle as CO
    cfg = dict(CO.TINY_CLIP_CONFIG)  # 4 heads of 32 (padded to 64), 2x2 patches
    oc, mc = _pair(cfg)
    g = torch.Generator("cpu").manual_seed(1)
    x = torch.randn(3, 3, 28, 28, generator=g).to(dev)
    with torch.no_grad():
        want = oc(x)
    got = mc(x).image_embeds
    assert got.shape == want.shape == (3, 64)
    assert rel(got, want) < 2e-2, rel(got, want)
    img = torch.rand(2, 3, 64, 96, generator=g).to(dev)
    for clamp in (False, True):
        with torch.no_grad():
            w2 = CO.encode_image(oc, img, clamp=clamp)
        g2 = mc.encode_image(img, clamp=clamp)
        assert g2.shape == w2.shape == (2, 1, 64)
        assert rel(g2, w2) < 2e-2, (clamp, rel(g2, w2))
    with pytest.raises(ValueError):
        mc(torch.zeros(1, 3, 30, 28))


def test_head_dim_80_padding_and_ragged_sequence():
    """ViT-H geometry in small: head_dim 80 (zero-padded to 128), 17 tokens (padded to 64)."""
    cfg = dict(hidden_size=320, intermediate_size=640, num_hidden_layers=2, num_attention_heads=4, image_size=56,
               patch_size=14, projection_dim=96, hidden_act="gelu", layer_norm_eps=1e-5)
    oc, mc = _pair(cfg, seed=2)
    g = torch.Generator("cpu").manual_seed(3)
    x = torch.randn(2, 3, 56, 56, generator=g).to(dev)
    with torch.no_grad():
        want = oc(x)
    got = mc(x).image_embeds
    assert rel(got, want) < 2e-2, rel(got, want)


def test_vit_h_full_size():
    """CLIP ViT-H/14 (632 M parameters, the SVD image encoder), 224 x 224, one image."""
    from oracle import clip_oracle as CO
    oc, mc = _pair({}, seed=4)
    g = torch.Generator("cpu").manual_seed(5)
    img = torch.rand(1, 3, 320, 512, generator=g).to(dev)
    with torch.no_grad():
        want = CO.encode_image(oc, img)
    got = mc.encode_image(img)
    assert got.shape == (1, 1, 1024)
    assert rel(got, want) < 3e-2, rel(got, want)


def test_pipeline_end_to_end_from_pixels():
    """f-1 + f-3 together: image and bbox frames in pixel space, CLIP + VAE encode, 25-step loop, VAE
    decode — everything on the sm_100a path — against the oracle chain."""
    import math
    from ctrlv_b200 import models, pipeline, vae, clip
    from oracle import clip_oracle as CO
    from oracle import sampling as S
    from oracle import svd_oracle as O
    from oracle import vae_oracle as V
    over = dict(O.TINY_CONFIG)
    ccfg = dict(CO.TINY_CLIP_CONFIG, projection_dim=over["cross_attention_dim"])
    torch.manual_seed(0)
    ou = O.UNetSpatioTemporalConditionModel(**over).to(dev).eval()
    oc = O.ControlNetModel(**over); O.randomize_zero_convs(oc); oc = oc.to(dev).eval()
    ov = V.AutoencoderKLTemporalDecoder(**V.TINY_VAE_CONFIG).to(dev).eval()
    oclip = CO.CLIPVisionModelWithProjection(**ccfg).to(dev).eval()
    pipe = pipeline.StableVideoControlPipeline(
        vae=vae.AutoencoderKLTemporalDecoder(state_dict=ov.state_dict(), **V.TINY_VAE_CONFIG),
        image_encoder=clip.CLIPVisionModelWithProjection(state_dict=oclip.state_dict(), **ccfg),
        unet=models.UNetSpatioTemporalConditionModel(state_dict=ou.state_dict(), **over),
        controlnet=models.ControlNetModel(state_dict=oc.state_dict(), **over))
    T, h, w, steps, aug = 3, 16, 16, 25, 0.02
    H, W = 2 * h, 2 * w
    g = torch.Generator("cpu").manual_seed(9)
    image = torch.rand(1, 3, H, W, generator=g)
    bbox = torch.rand(1, T, 3, H, W, generator=g) * 2 - 1
    lat0 = torch.randn(1, T, 4, h, w, generator=g)
    with torch.no_grad():
        emb = CO.encode_image(oclip, image.to(dev))
        noise = torch.randn(image.shape, generator=torch.Generator("cpu").manual_seed(13))
        il = ov.encode_mode((2 * image - 1 + aug * noise).to(dev)).unsqueeze(1).repeat(1, T, 1, 1, 1)
        ce = ov.encode_mode(bbox.flatten(0, 1).to(dev)).reshape(1, T, 4, h, w)
        inp = dict(latents=lat0.to(dev), image_latents=torch.cat([torch.zeros_like(il), il]),
                   image_embeddings=torch.cat([torch.zeros_like(emb), emb]),
                   cond_em=torch.cat([torch.zeros_like(ce), ce]),
                   added_time_ids=torch.tensor([[6.0, 127.0, aug]] * 2, device=dev),
                   guidance=torch.linspace(1.0, 3.0, T, device=dev))
        want = V.decode_latents(ov, S.sample_loop(ou, oc, inp, num_steps=steps), T, T)
        want = (want.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
    got = pipe(image=image, cond_images=bbox, height=H, width=W, num_frames=T, num_inference_steps=steps,
               latents=lat0.clone(), output_type="pt", noise_aug_strength=aug,
               generator=torch.Generator("cpu").manual_seed(13)).frames
    mse = float(((got.float() - want.float()) ** 2).mean())
    p = 10 * math.log10(1.0 / max(mse, 1e-30))
    assert got.shape == (1, T, 3, H, W) and p >= 40.0, p
    # PIL input takes the same route
    from PIL import Image
    pil = Image.fromarray((image[0].permute(1, 2, 0).numpy() * 255).round().astype("uint8"))
    got2 = pipe(image=pil, cond_images=bbox, height=H, width=W, num_frames=T, num_inference_steps=steps,
                latents=lat0.clone(), output_type="pil", noise_aug_strength=aug,
                generator=torch.Generator("cpu").manual_seed(13)).frames
    assert len(got2) == 1 and len(got2[0]) == T and got2[0][0].size == (W, H)


def test_pipeline_from_pretrained_directory(tmp_path):
    """tools/eval_video_controlnet.py:114-118: load unet/controlnet from a training output directory and the
    rest of the pipeline (vae, image_encoder) from a local Stable-Video-Diffusion directory."""
    from ctrlv_b200 import models, pipeline, vae, clip
    from oracle import clip_oracle as CO
    from oracle import svd_oracle as O
    from oracle import vae_oracle as V
    over = dict(O.TINY_CONFIG)
    ccfg = dict(CO.TINY_CLIP_CONFIG, projection_dim=over["cross_attention_dim"])
    svd, run = str(tmp_path / "svd"), str(tmp_path / "run")
    mv = vae.AutoencoderKLTemporalDecoder(seed=3, **V.TINY_VAE_CONFIG); mv.save_pretrained(svd, subfolder="vae")
    me = clip.CLIPVisionModelWithProjection(seed=4, **ccfg); me.save_pretrained(svd, subfolder="image_encoder")
    mu = models.UNetSpatioTemporalConditionModel(seed=5, **over); mu.save_pretrained(run, subfolder="unet")
    mc = models.ControlNetModel(seed=6, **over); mc.save_pretrained(run, subfolder="controlnet")
    ctrlnet = models.ControlNetModel.from_pretrained(run, subfolder="controlnet")
    unet = models.UNetSpatioTemporalConditionModel.from_pretrained(run, subfolder="unet")
    pipe = pipeline.StableVideoControlPipeline.from_pretrained(svd, controlnet=ctrlnet, unet=unet).to("cuda")
    pipe.set_progress_bar_config(disable=True)
    assert pipe.vae is not None and pipe.image_encoder is not None and pipe.unet is unet and pipe.controlnet is ctrlnet
    ref = pipeline.StableVideoControlPipeline(vae=mv, image_encoder=me, unet=mu, controlnet=mc)
    g = torch.Generator("cpu").manual_seed(21)
    T, h, w = 2, 8, 8
    image = torch.rand(1, 3, 2 * h, 2 * w, generator=g)
    bbox = torch.rand(1, T, 3, 2 * h, 2 * w, generator=g) * 2 - 1
    lat0 = torch.randn(1, T, 4, h, w, generator=g)
    kw = dict(image=image, cond_images=bbox, height=2 * h, width=2 * w, num_frames=T, num_inference_steps=4,
              output_type="pt", noise_aug_strength=0.02)
    a = pipe(latents=lat0.clone(), generator=torch.Generator("cpu").manual_seed(1), **kw).frames
    b = ref(latents=lat0.clone(), generator=torch.Generator("cpu").manual_seed(1), **kw).frames
    assert torch.equal(a, b)
    with pytest.raises(OSError):
        pipeline.StableVideoControlPipeline.from_pretrained(str(tmp_path / "missing"))
