"""GPU parity of the temporal VAE against the fp32 oracle restatement of
diffusers' AutoencoderKLTemporalDecoder: glue kernels, encoder `.mode()`, temporal decoder, and the
decoded-frame PSNR criterion of the north star (>= 40 dB)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
dev = "cuda"
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def psnr(a, b):
    """PSNR with the oracle's own dynamic range as the peak (random-init decoders are not in [-1, 1])."""
    a, b = a.float(), b.float()
    mse = float(((a - b) ** 2).mean())
    peak = float(b.max() - b.min())
    return 10.0 * math.log10(peak * peak / max(mse, 1e-30))


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


@pytest.mark.parametrize("frames,H,W,C,N", [(2, 8, 12, 64, 64), (3, 16, 16, 128, 128), (1, 6, 10, 64, 96)])
def test_conv3x3_stride2_asymmetric_pad(frames, H, W, C, N):
    from ctrlv_b200 import ops
    g = torch.Generator("cpu").manual_seed(0)
    x = torch.randn(frames, C, H, W, generator=g).to(dev)
    w = (torch.randn(N, C, 3, 3, generator=g) / (9 * C) ** 0.5).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    xb, wb = x.to(BF), w.to(BF)
    want = F.conv2d(F.pad(xb.float(), (0, 1, 0, 1)), wb.float(), b, stride=2)
    rows = xb.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    w9 = wb.permute(0, 2, 3, 1).reshape(N, -1).contiguous()
    got = ops.conv3x3_s2_pad01(rows, frames, H, W, w9, bias=b)
    got = got.float().reshape(frames, H // 2, W // 2, N).permute(0, 3, 1, 2)
    assert rel(got, want) < 5e-3


@pytest.mark.parametrize("M,N,pad", [(64, 64, 0), (130, 2560, 0), (77, 100, 28), (5, 7, 1)])
def test_softmax_rows(M, N, pad):
    from ctrlv_b200 import ops
    g = torch.Generator("cpu").manual_seed(1)
    s = (4.0 * torch.randn(M, N + pad, generator=g)).to(dev)
    out = torch.zeros(M, N + pad, dtype=BF, device=dev)
    ops.softmax_rows(s[:, :N], 0.37, out=out[:, :N])
    want = torch.softmax(0.37 * s[:, :N], dim=-1)
    assert float((out[:, :N].float() - want).abs().max()) < 4e-3
    assert float((out[:, :N].float().sum(-1) - 1).abs().max()) < 2e-2
    if pad:
        assert float(out[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("B,T,H,W", [(1, 1, 4, 4), (2, 5, 8, 6), (1, 14, 16, 16)])
def test_time_conv_out(B, T, H, W):
    from ctrlv_b200 import ops
    g = torch.Generator("cpu").manual_seed(2)
    x = torch.randn(B * T * H * W, 4, generator=g).to(dev)
    w = torch.randn(3, 3, 3, generator=g).to(dev)
    b = torch.randn(3, generator=g).to(dev)
    got = ops.time_conv_out(x, B, T, H, W, 3, w, b)
    x5 = x[:, :3].reshape(B, T, H, W, 3).permute(0, 4, 1, 2, 3)
    want = F.conv3d(x5, w.reshape(3, 3, 3, 1, 1), b, padding=(1, 0, 0)).permute(0, 2, 1, 3, 4).reshape(B * T, 3, H, W)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def _pair(over, seed=0):
    from ctrlv_b200 import vae
    from oracle import vae_oracle as V
    torch.manual_seed(seed)
    ov = V.AutoencoderKLTemporalDecoder(**over).to(dev).eval()
    # mix factors away from 0 so that the blend weights matter
    with torch.no_grad():
        for n, p in ov.named_parameters():
            if n.endswith("mix_factor"):
                p.fill_(0.3)
    mv = vae.AutoencoderKLTemporalDecoder(state_dict=ov.state_dict(), **over)
    return ov, mv


@pytest.fixture(scope="module")
def tiny_vae():
    from oracle import vae_oracle as V
    return _pair(dict(V.TINY_VAE_CONFIG))


@pytest.mark.parametrize("B,T,h,w", [(1, 4, 8, 8), (2, 3, 8, 16), (1, 1, 5, 9)])
def test_decoder_matches_oracle(tiny_vae, B, T, h, w):
    ov, mv = tiny_vae
    g = torch.Generator("cpu").manual_seed(3)
    z = torch.randn(B * T, 4, h, w, generator=g).to(dev)
    with torch.no_grad():
        want = ov.decode(z, num_frames=T)
    got = mv.decode(z, num_frames=T).sample
    assert got.shape == want.shape and got.dtype == torch.float32
    assert rel(got, want) < 2e-2 and psnr(got, want) > 40.0, (rel(got, want), psnr(got, want))


@pytest.mark.parametrize("N,H,W", [(3, 16, 16), (2, 32, 16)])
def test_encoder_mode_matches_oracle(tiny_vae, N, H, W):
    ov, mv = tiny_vae
    g = torch.Generator("cpu").manual_seed(4)
    x = torch.rand(N, 3, H, W, generator=g).to(dev) * 2 - 1
    with torch.no_grad():
        want = ov.encode_mode(x)
        mom = ov.encode_moments(x)
    dist = mv.encode(x).latent_dist
    assert dist.mode().shape == want.shape
    assert rel(dist.mode(), want) < 2e-2, rel(dist.mode(), want)
    assert rel(torch.cat([dist.mean, dist.logvar], 1), mom) < 2e-2


def test_decode_latents_chunking_and_errors(tiny_vae):
    from ctrlv_b200 import vae
    from oracle import vae_oracle as V
    ov, mv = tiny_vae
    g = torch.Generator("cpu").manual_seed(5)
    lat = torch.randn(1, 4, 4, 8, 8, generator=g).to(dev)
    for chunk in (4, 2, 3):
        with torch.no_grad():
            want = V.decode_latents(ov, lat, 4, chunk)
        got = vae.decode_latents(mv, lat, 4, chunk)
        assert got.shape == (1, 3, 4, 16, 16)  # the tiny config has two levels: x2
        assert psnr(got, want) > 40.0, (chunk, psnr(got, want))
    with pytest.raises(ValueError):
        mv.decode(torch.zeros(3, 4, 8, 8), num_frames=2)
    with pytest.raises(ValueError):
        mv.encode(torch.zeros(1, 3, 9, 8))


def test_full_architecture_decode_and_encode():
    """Full SVD VAE widths (128, 256, 512, 512), 97.7 M parameters, at a small resolution."""
    from oracle import vae_oracle as V
    ov, mv = _pair({}, seed=1)
    assert sum(p.numel() for p in ov.parameters()) == 97_742_847
    g = torch.Generator("cpu").manual_seed(6)
    T, h, w = 3, 8, 8
    z = torch.randn(T, 4, h, w, generator=g).to(dev)
    with torch.no_grad():
        want = ov.decode(z, num_frames=T)
    got = mv.decode(z, num_frames=T).sample
    assert got.shape == (T, 3, 8 * h, 8 * w)
    assert psnr(got, want) > 40.0 and rel(got, want) < 2e-2, (psnr(got, want), rel(got, want))
    x = torch.rand(2, 3, 64, 64, generator=g).to(dev) * 2 - 1
    with torch.no_grad():
        wz = ov.encode_mode(x)
    gz = mv.encode(x).latent_dist.mode()
    assert rel(gz, wz) < 2e-2, rel(gz, wz)


def test_pipeline_pixels_in_pixels_out_psnr():
    """North-star parity criterion: decoded frames after the 25-step loop, new path vs
    the fp32 oracle chain (VAE encode -> ControlNet+UNet Euler loop -> temporal VAE decode), PSNR >= 40 dB.
    Pixel-space bbox frames and conditioning image go in, frames come out (pipeline_video_control.py:84,235,346)."""
    from ctrlv_b200 import models, pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    from oracle import vae_oracle as V
    ov, mv = _pair(dict(V.TINY_VAE_CONFIG))           # two levels: pixels = 2 x latent
    over = dict(O.TINY_CONFIG)
    torch.manual_seed(0)
    ou = O.UNetSpatioTemporalConditionModel(**over).to(dev).eval()
    oc = O.ControlNetModel(**over)
    O.randomize_zero_convs(oc)
    oc = oc.to(dev).eval()
    mu = models.UNetSpatioTemporalConditionModel(state_dict=ou.state_dict(), **over)
    mc = models.ControlNetModel(state_dict=oc.state_dict(), **over)
    T, h, w, steps, aug = 4, 16, 16, 25, 0.02
    H, W = 2 * h, 2 * w
    g = torch.Generator("cpu").manual_seed(7)
    image = torch.rand(1, 3, H, W, generator=g)
    bbox = torch.rand(1, T, 3, H, W, generator=g) * 2 - 1
    emb = torch.randn(1, 1, over["cross_attention_dim"], generator=g)
    lat0 = torch.randn(1, T, 4, h, w, generator=g)
    # oracle chain
    with torch.no_grad():
        noise = torch.randn(image.shape, generator=torch.Generator("cpu").manual_seed(11))
        il = ov.encode_mode((2 * image - 1 + aug * noise).to(dev))
        ce = ov.encode_mode(bbox.flatten(0, 1).to(dev)).reshape(1, T, 4, h, w)
        ilr = il.unsqueeze(1).repeat(1, T, 1, 1, 1)
        inp = dict(latents=lat0.to(dev), image_latents=torch.cat([torch.zeros_like(ilr), ilr]),
                   image_embeddings=torch.cat([torch.zeros_like(emb), emb]).to(dev),
                   cond_em=torch.cat([torch.zeros_like(ce), ce]),
                   added_time_ids=torch.tensor([[6.0, 127.0, aug]] * 2, device=dev),
                   guidance=torch.linspace(1.0, 3.0, T, device=dev))
        lat_o = S.sample_loop(ou, oc, inp, num_steps=steps)
        want = V.decode_latents(ov, lat_o, T, T)
        want = (want.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
    # new path: pixels in, pixels out
    pipe = pipeline.StableVideoControlPipeline(vae=mv, unet=mu, controlnet=mc)
    assert pipe.vae_scale_factor == 2
    out = pipe(image=image, cond_images=bbox, height=H, width=W, num_frames=T, num_inference_steps=steps,
               latents=lat0.clone(), output_type="pt", image_embeddings=emb, noise_aug_strength=aug,
               generator=torch.Generator("cpu").manual_seed(11))
    got = out.frames
    assert got.shape == (1, T, 3, H, W) and float(got.min()) >= 0 and float(got.max()) <= 1
    mse = float(((got.float() - want.float()) ** 2).mean())
    p = 10 * math.log10(1.0 / max(mse, 1e-30))
    assert p >= 40.0, p
    arr = pipe(image=image, cond_images=bbox, height=H, width=W, num_frames=T, num_inference_steps=steps,
               latents=lat0.clone(), output_type="np", image_embeddings=emb, noise_aug_strength=aug,
               generator=torch.Generator("cpu").manual_seed(11), decode_chunk_size=2).frames
    assert arr.shape == (1, T, H, W, 3)
    lat = pipe(image=image, cond_images=bbox, height=H, width=W, num_frames=T, num_inference_steps=steps,
               latents=lat0.clone(), output_type="latent", image_embeddings=emb, noise_aug_strength=aug,
               generator=torch.Generator("cpu").manual_seed(11)).frames
    assert rel(lat, lat_o) < 2e-2, rel(lat, lat_o)
