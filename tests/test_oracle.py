"""CPU tests that pin the oracle restatement: parameter counts, key set,
scheduler / embedding known answers, algebraic identities, and the committed golden fixture."""
import math
import os

import pytest
import torch

from oracle import sampling as S
from oracle import svd_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_step.pt")


def test_param_counts_exact():
    with torch.device("meta"):
        u = O.UNetSpatioTemporalConditionModel()
        c = O.ControlNetModel()
    assert sum(p.numel() for p in u.parameters()) == 1_524_623_082
    assert sum(p.numel() for p in c.parameters()) == 680_946_897
    # controlnet.py:148-185: 1 + (2+1)*3 + 2 = 12 zero-convs + 1 mid
    assert len(c.controlnet_down_blocks) == 12


def test_state_dict_key_names():
    with torch.device("meta"):
        u = O.UNetSpatioTemporalConditionModel()
        c = O.ControlNetModel()
    uk, ck = set(u.state_dict()), set(c.state_dict())
    for k in ["conv_in.weight", "time_embedding.linear_1.weight", "add_embedding.linear_2.bias",
              "down_blocks.0.resnets.0.spatial_res_block.conv1.weight",
              "down_blocks.0.resnets.0.temporal_res_block.conv1.weight",
              "down_blocks.0.resnets.0.time_mixer.mix_factor",
              "down_blocks.1.resnets.0.spatial_res_block.conv_shortcut.weight",
              "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_out.0.bias",
              "down_blocks.0.attentions.0.temporal_transformer_blocks.0.ff_in.net.0.proj.weight",
              "down_blocks.0.attentions.0.time_pos_embed.linear_1.weight",
              "down_blocks.0.downsamplers.0.conv.weight", "mid_block.attentions.0.proj_out.weight",
              "up_blocks.0.upsamplers.0.conv.weight", "up_blocks.3.attentions.2.norm.weight",
              "conv_norm_out.weight", "conv_out.bias"]:
        assert k in uk, k
    for k in ["control_conv_in.weight", "controlnet_down_blocks.11.bias", "controlnet_mid_block.weight"]:
        assert k in ck and k not in uk, k
    assert not any(k.startswith("up_blocks") or k.startswith("conv_out") for k in ck)
    # from_unet copies exactly the shared keys (controlnet.py:214-220)
    assert (ck & uk) == {k for k in ck if not k.startswith("control")}


def test_scheduler_known_answers():
    # SURVEY.md A.9 (EulerDiscreteScheduler, SVD config, N = 25)
    s = S.EulerDiscreteSchedulerOracle()
    s.set_timesteps(25)
    sig, t = s.sigmas, s.timesteps
    assert sig.shape == (26,) and sig[-1] == 0
    for got, want in zip(sig[:4].tolist(), [700.0, 545.729248, 421.569122, 322.453674]):
        assert abs(got - want) / want < 1e-6
    for got, want in zip(sig[22:25].tolist(), [0.0248025805, 0.00788249541, 0.00200000009]):
        assert abs(got - want) / want < 1e-5
    for got, want in zip(t[:4].tolist(), [1.63777006, 1.57553077, 1.51099598, 1.44398987]):
        assert abs(got - want) < 1e-6
    for got, want in zip(t[22:25].tolist(), [-0.924201906, -1.21077764, -1.55365205]):
        assert abs(got - want) < 1e-5
    assert abs(float(s.init_noise_sigma) - 700.000732) < 1e-3


def test_euler_step_matches_training_side_formulas():
    # tools/train_video_controlnet.py:468-471: c_out = -sigma/sqrt(sigma^2+1), c_skip = 1/(sigma^2+1)
    s = S.EulerDiscreteSchedulerOracle()
    s.set_timesteps(25)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 2, 4, 4, 4, generator=g) * 50
    v = torch.randn(1, 2, 4, 4, 4, generator=g)
    i = 7
    s.step_index = i
    out = s.step(v, s.timesteps[i], x)
    sigma, nxt = float(s.sigmas[i]), float(s.sigmas[i + 1])
    x0 = (-sigma / math.sqrt(sigma ** 2 + 1)) * v + x / (sigma ** 2 + 1)
    ref = x + (x - x0) / sigma * (nxt - sigma)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    assert torch.allclose(s.scale_model_input(x, s.timesteps[i + 1]), x / math.sqrt(nxt ** 2 + 1))


def test_timesteps_embedding_known_answer():
    # SURVEY.md A.8
    e = O.Timesteps(320, True, 0)(torch.tensor([1.63777006]))
    assert torch.allclose(e[0, :3], torch.tensor([-0.0669237, 0.0246392, 0.1109036]), atol=1e-6)
    assert torch.allclose(e[0, 160:163], torch.tensor([0.9977581, 0.9996964, 0.9938312]), atol=1e-6)


def _tiny(seed=0):
    torch.manual_seed(seed)
    u = O.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG).eval()
    c = O.ControlNetModel(**O.TINY_CONFIG).eval()
    return u, c


def _tiny_inputs(T=2, h=8, w=8):
    return S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"])


def test_zero_convs_make_controlnet_a_noop_at_init():
    # controlnet.py:148-185: zero-initialised 1x1 convs => residuals are exactly 0
    u, c = _tiny()
    inp = _tiny_inputs()
    x = torch.cat([torch.cat([inp["latents"]] * 2), inp["image_latents"]], dim=2)
    t = torch.tensor(1.2)
    with torch.no_grad():
        d, m = c(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        assert all(float(r.abs().max()) == 0.0 for r in d) and float(m.abs().max()) == 0.0
        y0 = u(x, t, inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
        y1 = u(x, t, inp["image_embeddings"], inp["added_time_ids"], d, m, return_dict=False)[0]
    assert torch.equal(y0, y1)
    shapes = [tuple(r.shape[1:]) for r in d]
    assert shapes == [(64, 8, 8)] * 3 + [(64, 4, 4)] + [(128, 4, 4)] * 2 + [(128, 2, 2)] + [(256, 2, 2)] * 2 + [(256, 1, 1)] * 3


def test_single_token_cross_attention_is_a_vector():
    # softmax over one key is 1  =>  attn2(x, ctx) = to_out(to_v(ctx))
    torch.manual_seed(1)
    a = O.Attention(128, 96, 2, 64)
    x = torch.randn(3, 10, 128)
    ctx = torch.randn(3, 1, 96)
    with torch.no_grad():
        got = a(x, ctx)
        want = a.to_out[0](a.to_v(ctx)).expand(3, 10, 128)
    assert torch.allclose(got, want, atol=1e-6)


def test_time_context_order_only_matters_for_batch_gt_1():
    torch.manual_seed(2)
    m = O.TransformerSpatioTemporalModel(2, 64, 128, cross_attention_dim=96, time_context_order="s_major").eval()
    m2 = O.TransformerSpatioTemporalModel(2, 64, 128, cross_attention_dim=96, time_context_order="b_major").eval()
    m2.load_state_dict(m.state_dict())
    with torch.no_grad():
        for B, same in ((1, True), (2, False)):
            T = 3
            x = torch.randn(B * T, 128, 4, 4)
            ehs = torch.randn(B, 1, 96).repeat_interleave(T, 0)
            ind = torch.zeros(B, T)
            assert torch.allclose(m(x, ehs, ind), m2(x, ehs, ind), atol=1e-5) == same


def test_from_unet_copies_shared_weights():
    u, _ = _tiny()
    c = O.ControlNetModel.from_unet(u)
    usd, csd = u.state_dict(), c.state_dict()
    for k in csd:
        if k in usd:
            assert torch.equal(csd[k], usd[k]), k
    assert float(c.controlnet_mid_block.weight.abs().max()) == 0.0


def test_golden_fixture():
    from tests.golden import make_golden
    want = torch.load(GOLDEN)
    got = make_golden.run()
    assert torch.equal(got["sigmas"], want["sigmas"]) and torch.equal(got["timesteps"], want["timesteps"])
    assert abs(got["unet_param_checksum"] - want["unet_param_checksum"]) < 1e-3 * abs(want["unet_param_checksum"]) + 1e-3
    for k in ("noise_pred", "latents_after_step0", "ctrl_down_norms", "ctrl_mid"):
        err = (got[k] - want[k]).norm() / want[k].norm()
        assert err < 1e-4, (k, float(err))


def test_vae_oracle_structure_pins():
    """f-1 oracle (AutoencoderKLTemporalDecoder restatement): parameter counts of the SVD VAE config
    (encoder = the Stable Diffusion VAE encoder, 34,163,592; whole module 97,742,847) and the
    diffusers key families a real checkpoint carries."""
    from oracle import vae_oracle as V
    with torch.device("meta"):
        m = V.AutoencoderKLTemporalDecoder()
    assert sum(p.numel() for p in m.encoder.parameters()) == 34_163_592
    assert sum(p.numel() for p in m.quant_conv.parameters()) == 72
    assert sum(p.numel() for p in m.parameters()) == 97_742_847
    keys = set(m.state_dict())
    for k in ("encoder.down_blocks.0.downsamplers.0.conv.weight", "encoder.mid_block.attentions.0.group_norm.weight",
              "encoder.mid_block.attentions.0.to_out.0.bias", "decoder.mid_block.resnets.1.temporal_res_block.conv2.weight",
              "decoder.up_blocks.3.resnets.2.time_mixer.mix_factor", "decoder.up_blocks.2.upsamplers.0.conv.bias",
              "decoder.time_conv_out.weight", "quant_conv.bias"):
        assert k in keys, k
    assert not any(k.startswith("post_quant_conv") or "time_emb_proj" in k for k in keys)
    assert "decoder.up_blocks.3.upsamplers.0.conv.weight" not in keys


def test_vae_oracle_algebra():
    """Temporal layers only mix frames of one clip; with sigmoid(mix)=0 blend the decoder is per-frame."""
    from oracle import vae_oracle as V
    torch.manual_seed(0)
    m = V.AutoencoderKLTemporalDecoder(**V.TINY_VAE_CONFIG).eval()
    z = torch.randn(4, 4, 6, 6)
    with torch.no_grad():
        a = m.decode(z, num_frames=2)                       # two clips of two frames
        b = torch.cat([m.decode(z[:2], 2), m.decode(z[2:], 2)])
        assert torch.allclose(a, b, atol=1e-5)
        x = torch.rand(2, 3, 16, 16) * 2 - 1
        assert m.encode_mode(x).shape == (2, 4, 8, 8)
        f = V.decode_latents(m, z.reshape(1, 4, 4, 6, 6), 4, 2)
        assert f.shape == (1, 3, 4, 12, 12)
        # switch_spatial_to_temporal_mix: alpha = 1 - sigmoid(mix); mix -> +inf keeps only the temporal branch
        blk = m.decoder.mid_block.resnets[0]
        blk.time_mixer.mix_factor.data.fill_(-30.0)         # alpha -> 1: spatial branch only
        h = torch.randn(2, V.TINY_VAE_CONFIG["block_out_channels"][-1], 4, 4)
        ioi = torch.zeros(1, 2)
        assert torch.allclose(blk(h, None, ioi), blk.spatial_res_block(h, None), atol=1e-5)


def test_resize_oracle_matches_reference_golden():
    """PINNED: golden outputs of the reference's own `_resize_with_antialiasing`
    (tests/golden/make_resize_golden.py executes image_encoder.py:184-290 unmodified)."""
    from oracle import clip_oracle as CO
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "resize_antialias.pt"))
    assert len(cases) == 4
    for c in cases:
        got = CO.resize_with_antialiasing(c["input"], c["size"])
        assert got.shape == c["output"].shape
        assert torch.allclose(got, c["output"], atol=2e-6, rtol=1e-5), float((got - c["output"]).abs().max())


def test_clip_oracle_matches_transformers():
    """PINNED against the installed `transformers` CLIPVisionModelWithProjection (the reference's
    image_encoder class, pipeline_video_control.py:30): same state dict -> same image_embeds."""
    tf = pytest.importorskip("transformers")
    from oracle import clip_oracle as CO
    for act in ("gelu", "quick_gelu"):
        cfg = dict(CO.TINY_CLIP_CONFIG, hidden_act=act)
        hf_cfg = tf.CLIPVisionConfig(**cfg)
        torch.manual_seed(0)
        hf = tf.CLIPVisionModelWithProjection(hf_cfg).eval()
        mine = CO.CLIPVisionModelWithProjection(**cfg).eval()
        sd = {k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
        assert set(sd) == set(mine.state_dict()), set(sd) ^ set(mine.state_dict())
        mine.load_state_dict(sd)
        x = torch.randn(2, 3, cfg["image_size"], cfg["image_size"])
        with torch.no_grad():
            want = hf(pixel_values=x).image_embeds
            got = mine(x)
        assert torch.allclose(got, want, atol=1e-5, rtol=1e-4), float((got - want).abs().max())
    with torch.device("meta"):
        full = CO.CLIPVisionModelWithProjection()
    assert sum(p.numel() for p in full.parameters()) == 632_076_800  # ViT-H/14 vision tower + projection
