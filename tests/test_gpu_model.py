"""GPU parity of the drop-in models, the fused denoise step and the pipeline against the fp32
oracle (north_star tolerance: per-step latent relative L2 <= 1e-2 in bf16)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _build(order, over):
    from ctrlv_b200 import models
    from oracle import svd_oracle as O
    torch.manual_seed(0)
    ou = O.UNetSpatioTemporalConditionModel(time_context_order=order, **over)
    oc = O.ControlNetModel(time_context_order=order, **over)
    O.randomize_zero_convs(oc)
    ou, oc = ou.to(dev).eval(), oc.to(dev).eval()
    mu = models.UNetSpatioTemporalConditionModel(state_dict=ou.state_dict(), time_context_order=order, **over)
    mc = models.ControlNetModel(state_dict=oc.state_dict(), time_context_order=order, **over)
    return ou, oc, mu, mc


@pytest.fixture(scope="module")
def tiny():
    from oracle import svd_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _build("s_major", dict(O.TINY_CONFIG))


def _inputs(T, h, w, sigma):
    from oracle import svd_oracle as O
    from oracle import sampling as S
    inp = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], device=dev)
    x = torch.cat([inp["latents"] * (sigma ** 2 + 1) ** 0.5] * 2) / (sigma ** 2 + 1) ** 0.5
    return inp, torch.cat([x, inp["image_latents"]], dim=2)


@pytest.mark.parametrize("T,h,w", [(4, 16, 16), (3, 8, 24), (1, 8, 8), (25, 8, 16), (2, 40, 8)])
def test_controlnet_and_unet_forward(tiny, T, h, w):
    ou, oc, mu, mc = tiny
    inp, x = _inputs(T, h, w, 15.59)
    t = torch.tensor(0.6866)
    with torch.no_grad():
        od, om = oc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"],
                    conditioning_scale=0.8, return_dict=False)
        oy = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], od, om, return_dict=False)[0]
        oy0 = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
    out = mc(x, timestep=t.to(dev), encoder_hidden_states=inp["image_embeddings"],
             added_time_ids=inp["added_time_ids"], control_cond=inp["cond_em"], conditioning_scale=0.8)
    md, mm = out.down_block_res_samples, out.mid_block_res_sample
    assert len(md) == 12 and [tuple(a.shape) for a in md] == [tuple(b.shape) for b in od]
    assert tuple(mm.shape) == tuple(om.shape)
    my = mu(sample=x, timestep=t.to(dev), encoder_hidden_states=inp["image_embeddings"],
            added_time_ids=inp["added_time_ids"], down_block_additional_residuals=md,
            mid_block_additional_residuals=mm, return_dict=False)[0]
    my0 = mu(x, torch.stack([t, t]).to(dev), inp["image_embeddings"], inp["added_time_ids"]).sample  # [B] timestep
    torch.cuda.synchronize()
    assert max(rel(a, b) for a, b in zip(md, od)) < 2.5e-2 and rel(mm, om) < 2.5e-2
    assert rel(my, oy) < 2e-2 and rel(my0, oy0) < 2e-2
    assert rel(oy, oy0) > 5e-2  # the ControlNet path carries signal in this test


def test_residuals_in_reference_layout_are_accepted(tiny):
    """The UNet accepts residuals as plain contiguous NCHW fp32 tensors (what a reference
    ControlNet would hand over), not only this package's channels-last views."""
    ou, oc, mu, mc = tiny
    inp, x = _inputs(2, 8, 8, 3.0)
    t = torch.tensor(0.27, device=dev)
    md, mm = mc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
    a = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], md, mm, return_dict=False)[0]
    b = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], [d.float().contiguous() for d in md],
           mm.float().contiguous(), return_dict=False)[0]
    assert torch.equal(a, b)


def test_zero_init_controlnet_is_a_noop_and_from_unet(tiny):
    """controlnet.py:149-184: the zero-convs are `zero_module`s, so a fresh `from_unet` ControlNet (and a
    fresh constructor) leaves the UNet untouched — no hand-zeroing of weights."""
    from ctrlv_b200 import models
    from oracle import svd_oracle as O
    ou, oc, mu, mc = tiny
    inp, x = _inputs(2, 8, 8, 3.0)
    t = torch.tensor(0.27, device=dev)
    y0 = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
    for c0 in (models.ControlNetModel.from_unet(mu), models.ControlNetModel(seed=3, **O.TINY_CONFIG)):
        d, m = c0(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        assert all(float(r.abs().max()) == 0.0 for r in d) and float(m.abs().max()) == 0.0
        y1 = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], d, m, return_dict=False)[0]
        assert torch.equal(y0, y1)
    c1 = models.ControlNetModel.from_unet(mu)
    su, sc = mu.state_dict(), c1.state_dict()
    assert all(torch.equal(sc[k].to(dev), su[k].to(dev)) for k in sc if k in su)  # encoder weights copied
    c2 = models.ControlNetModel(seed=3, zero_conv_std=0.02, **O.TINY_CONFIG)      # opt-in: residual path carries signal
    d, m = c2(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
    assert float(m.abs().max()) > 0.0


def test_conditioning_scale_zero_switches_the_controlnet_off(tiny):
    """controlnet.py:343-344 multiplies the residuals by `conditioning_scale`: 0 gives all-zero residuals and
    the UNet output of the plain (no-ControlNet) forward — the no-control ablation."""
    from ctrlv_b200 import pipeline
    from oracle import sampling as S
    ou, oc, mu, mc = tiny
    inp, x = _inputs(2, 8, 8, 3.0)
    t = torch.tensor(0.27, device=dev)
    d, m = mc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"],
              conditioning_scale=0.0, return_dict=False)
    assert all(float(r.abs().max()) == 0.0 for r in d) and float(m.abs().max()) == 0.0
    d1, m1 = mc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
    assert float(m1.abs().max()) > 0.0
    y0 = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
    y = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], d, m, return_dict=False)[0]
    assert torch.equal(y, y0)
    # the pipeline's control_condition_scale = 0 equals a pipeline without a ControlNet
    T, h, w = 2, 8, 8
    kw = dict(cond_images=inp["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T, num_inference_steps=3,
              latents=inp["latents"].clone(), output_type="latent", image_embeddings=inp["image_embeds_cond"],
              image_latents=inp["image_latents_cond"])
    a = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)(control_condition_scale=0.0, **kw).frames
    b = pipeline.StableVideoControlPipeline(unet=mu, controlnet=None)(**kw).frames
    c = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)(control_condition_scale=1.0, **kw).frames
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_reloading_weights_invalidates_cached_graphs(tiny):
    """A captured step graph holds raw pointers into the packed weights; `load_state_dict` repacks, so the
    pipeline must not replay a graph captured before it."""
    from ctrlv_b200 import models, pipeline
    from oracle import svd_oracle as O
    ou, oc, mu, mc = tiny
    mu2 = models.UNetSpatioTemporalConditionModel(state_dict=mu.state_dict(), **O.TINY_CONFIG)
    inp, _ = _inputs(2, 8, 8, 3.0)
    kw = dict(cond_images=inp["cond_em_cond"], height=64, width=64, num_frames=2, num_inference_steps=3,
              latents=inp["latents"].clone(), output_type="latent", image_embeddings=inp["image_embeds_cond"],
              image_latents=inp["image_latents_cond"])
    pipe = pipeline.StableVideoControlPipeline(unet=mu2, controlnet=mc)
    a = pipe(**kw).frames
    sd = {k: (v * 1.5 if k == "conv_out.weight" else v) for k, v in mu2.state_dict().items()}
    mu2.load_state_dict(sd)
    junk = [torch.randn(1 << 20, device=dev) for _ in range(8)]  # recycle the freed blocks
    b = pipe(**kw).frames
    fresh = pipeline.StableVideoControlPipeline(unet=mu2, controlnet=mc)(**kw).frames
    assert torch.equal(b, fresh) and not torch.equal(a, b)
    assert len(pipe._steps) == 2
    pipe.max_cached_steps = 2
    pipe(**dict(kw, num_inference_steps=4)); pipe(**dict(kw, num_inference_steps=5))
    assert len(pipe._steps) == 2  # LRU bound
    del junk


def test_num_videos_per_prompt_duplication_orders():
    """pipeline_video_control.py:90 tiles the bbox-frame latents; diffusers-0.27.2 `_encode_vae_image` tiles the
    image latents and `_encode_image` interleaves the embeddings."""
    from ctrlv_b200.pipeline import StableVideoControlPipeline as P
    emb = torch.arange(2.0).view(2, 1, 1).repeat(1, 1, 4)
    il = torch.arange(2.0).view(2, 1, 1, 1).repeat(1, 4, 2, 2)
    e, l = P._duplicate_conditioning(emb, il, 2, 3)
    assert e[:, 0].tolist() == [0, 0, 0, 1, 1, 1] and l[:, 0, 0, 0].tolist() == [0, 1, 0, 1, 0, 1]
    pipe = P.__new__(P)
    pipe.vae = None
    cond = torch.arange(2.0).view(2, 1, 1, 1, 1).repeat(1, 2, 4, 2, 2)
    ce = pipe._encode_vae_condition(cond, 3, True)
    assert ce[6:, 0, 0, 0, 0].tolist() == l[:, 0, 0, 0].tolist() and float(ce[:6].abs().max()) == 0.0


@pytest.mark.parametrize("order", ["s_major", "b_major"])
def test_pipeline_loop_matches_oracle_loop(order):
    from ctrlv_b200 import pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    ou, oc, mu, mc = _build(order, dict(O.TINY_CONFIG))
    T, h, w, steps = 4, 16, 16, 25  # the reference's default schedule (pipeline_video_control.py:112)
    inp = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], device=dev)
    trace, mtrace = [], []
    with torch.no_grad():
        S.sample_loop(ou, oc, inp, num_steps=steps, trace=trace)
    pipe = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)
    res = {}
    for use_graph in (False, True):
        mtrace.clear()
        out = pipe(cond_images=inp["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T,
                   num_inference_steps=steps, latents=inp["latents"].clone(), output_type="latent",
                   image_embeddings=inp["image_embeds_cond"], image_latents=inp["image_latents_cond"],
                   use_graph=use_graph,
                   callback_on_step_end=lambda p, i, t, kw: mtrace.append(kw["latents"].clone()) or {})
        torch.cuda.synchronize()
        res[use_graph] = out.frames.clone()
        # teacher-forced per-step error: restart each step from the oracle's state
        st = next(s for s in pipe._steps.values() if s.use_graph == use_graph)
        sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(steps)
        prevs = [inp["latents"] * sch.init_noise_sigma] + trace[:-1]
        for i in range(steps):
            st.latents.copy_(prevs[i]); st.step(i)
            assert rel(st.latents, trace[i]) < 1e-2, (order, use_graph, i)
    assert torch.equal(res[False], res[True])  # graph replay == eager launches, bit for bit


def test_bbox_predictor_pipeline_matches_oracle_loop(tiny):
    """VideoDiffusionPipeline (plain SVD sampler + conditioning-frame overwrite,
    pipeline_video_diffusion.py:196-293) against the oracle loop, teacher-forced per step."""
    from ctrlv_b200 import pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    ou, _, mu, _ = tiny
    T, h, w, steps = 5, 16, 8, 25
    inp = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], device=dev)
    trace = []
    with torch.no_grad():
        S.sample_loop_bbox_predictor(ou, inp, num_steps=steps, num_cond_bbox_frames=2, trace=trace)
    pipe = pipeline.VideoDiffusionPipeline(unet=mu)
    out = pipe(bbox_images=inp["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T,
               num_inference_steps=steps, latents=inp["latents"].clone(), output_type="latent",
               image_embeddings=inp["image_embeds_cond"], image_latents=inp["image_latents_cond"],
               num_cond_bbox_frames=2)
    assert tuple(out.frames.shape) == (1, T, 4, h, w) and torch.isfinite(out.frames).all()
    st = next(iter(pipe._steps.values()))
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(steps)
    prevs = [inp["latents"] * sch.init_noise_sigma] + trace[:-1]
    for i in range(steps):
        st.latents.copy_(prevs[i]); st.step(i)
        assert rel(st.latents, trace[i]) < 1e-2, i
    # without bbox frames the pipeline is the stock SVD sampler: image latents repeated over T
    out2 = pipe(height=h * 8, width=w * 8, num_frames=T, num_inference_steps=steps,
                latents=inp["latents"].clone(), output_type="latent",
                image_embeddings=inp["image_embeds_cond"], image_latents=inp["image_latents_cond"])
    assert not torch.equal(out2.frames, out.frames)


def test_checkpoint_roundtrip_from_pretrained(tiny, tmp_path):
    """save_pretrained -> from_pretrained (diffusers directory layout, the call of
    tools/eval_video_controlnet.py:114-118) reproduces the networks bit for bit."""
    from ctrlv_b200 import models
    _, _, mu, mc = tiny
    mu.save_pretrained(str(tmp_path), subfolder="unet")
    mc.save_pretrained(str(tmp_path), subfolder="controlnet", max_shard_bytes=1 << 20)
    mu2 = models.UNetSpatioTemporalConditionModel.from_pretrained(str(tmp_path), subfolder="unet")
    mc2 = models.ControlNetModel.from_pretrained(str(tmp_path), subfolder="controlnet")
    assert mu2.cfg == mu.cfg and mc2.cfg == mc.cfg
    for a, b in ((mu, mu2), (mc, mc2)):
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    inp, x = _inputs(3, 8, 16, 3.0)
    t = torch.tensor(0.27, device=dev)
    kw = dict(timestep=t, encoder_hidden_states=inp["image_embeddings"], added_time_ids=inp["added_time_ids"])
    d1, m1 = mc(x, control_cond=inp["cond_em"], return_dict=False, **kw)
    d2, m2 = mc2(x, control_cond=inp["cond_em"], return_dict=False, **kw)
    assert torch.equal(m1, m2) and all(torch.equal(a, b) for a, b in zip(d1, d2))
    y1 = mu(sample=x, down_block_additional_residuals=d1, mid_block_additional_residuals=m1, return_dict=False, **kw)[0]
    y2 = mu2(sample=x, down_block_additional_residuals=d2, mid_block_additional_residuals=m2, return_dict=False, **kw)[0]
    assert torch.equal(y1, y2)
    with pytest.raises(OSError):
        models.ControlNetModel.from_pretrained(str(tmp_path), subfolder="nope")


def test_pipeline_errors():
    from ctrlv_b200 import pipeline, models
    from oracle import svd_oracle as O
    mu = models.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG)
    pipe = pipeline.StableVideoControlPipeline(unet=mu, controlnet=None)
    with pytest.raises(ValueError):
        pipe(cond_images=None, height=64, width=64)
    with pytest.raises(ValueError):
        pipe(cond_images=torch.zeros(1, 2, 4, 8, 8), height=65, width=64)
    with pytest.raises(NotImplementedError):
        pipe(cond_images=torch.zeros(1, 2, 4, 8, 8), height=64, width=64, num_frames=2)


@pytest.mark.parametrize("T,h,w", [(14, 40, 64)])
def test_full_size_single_step(T, h, w):
    """BASELINE config 1/2 shapes with the full SVD architecture: one CFG step, noise_pred and
    teacher-forced latent error vs the fp32 oracle (run on the GPU in fp32)."""
    from ctrlv_b200 import models, pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = dict(models.SVD_CONFIG)
    sd_u = models.random_state_dict(cfg, False, seed=0, dtype=torch.float32)
    sd_c = models.random_state_dict(cfg, True, seed=1, dtype=torch.float32, zero_conv_std=0.02)
    with torch.device("meta"):
        ou = O.UNetSpatioTemporalConditionModel(); oc = O.ControlNetModel()
    ou.load_state_dict(sd_u, assign=True); oc.load_state_dict(sd_c, assign=True)
    mu = models.UNetSpatioTemporalConditionModel(state_dict=sd_u)
    mc = models.ControlNetModel(state_dict=sd_c)
    inp = S.make_inputs(T=T, h=h, w=w, device=dev)
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
    st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=False)
    st.set_schedule(sch.sigmas, sch.timesteps)
    st.image_latents.copy_(inp["image_latents"]); st.cond_em.copy_(inp["cond_em"])
    st.ehs.copy_(inp["image_embeddings"].reshape(2, -1)); st.added_time_ids.copy_(inp["added_time_ids"])
    st.guidance.copy_(inp["guidance"])
    gs = inp["guidance"].view(1, -1, 1, 1, 1)
    for i in (0, 17):
        sigma = sch.sigmas[i]
        lat = inp["latents"] * float((sigma ** 2 + 1) ** 0.5)
        sch.step_index = i
        with torch.no_grad():
            want, noise = S.denoise_step(ou, oc, sch, lat, sch.timesteps[i], inp["image_latents"],
                                         inp["image_embeddings"], inp["added_time_ids"], inp["cond_em"], gs,
                                         return_noise=True)
        st.latents.copy_(lat)
        st.step(i)
        torch.cuda.synchronize()
        got_noise = st.noise.view(2, T, h, w, 4).permute(0, 1, 4, 2, 3)
        assert rel(got_noise, noise) < 2e-2, i          # model output (bf16 path vs fp32 oracle)
        assert rel(st.latents, want) < 1e-2, i          # north_star: per-step latent rel L2 <= 1e-2


@pytest.mark.parametrize("order,T,h,w,exact", [("s_major", 4, 16, 32, True), ("b_major", 4, 16, 32, True),
                                               ("s_major", 3, 8, 24, False)])
def test_cfg_branch_sharded_step_equals_batched_step(order, T, h, w, exact):
    """uncond and cond halves computed by two branch-sharded DenoiseSteps (here on one
    GPU, the exchange being a copy) reproduce the batched CFG step, including the diffusers-0.27.2
    time_context coupling of the two branches.  When no 32-row warp straddles two samples (T*S a
    multiple of 32 at every level) the result is bit-identical; in the 3x8x24 case (odd S = 3 at the
    mid level: exercises the context-table rotation) a straddling warp adds bias and time embedding in
    a different order, which flips single bf16 roundings."""
    from ctrlv_b200 import pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    _, _, mu, mc = _build(order, dict(O.TINY_CONFIG))
    steps = 25
    inp = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], device=dev)
    sch = pipeline.EulerDiscreteScheduler().set_timesteps(steps)

    def fill(st):
        st.set_schedule(sch.sigmas, sch.timesteps)
        st.image_latents.copy_(inp["image_latents"]); st.cond_em.copy_(inp["cond_em"])
        st.ehs.copy_(inp["image_embeddings"].reshape(2, -1)); st.added_time_ids.copy_(inp["added_time_ids"])
        st.guidance.copy_(inp["guidance"])
        st.capture()
        st.latents.copy_(inp["latents"] * sch.init_noise_sigma)

    full = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True)
    halves = [pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, cfg_branch=br) for br in (0, 1)]
    M = halves[0].noise_local.shape[0]

    def make_exchange(me):
        other = halves[1 - me]
        return lambda st: st.noise[(1 - me) * M:(2 - me) * M].copy_(other.noise_local)

    for br in (0, 1):
        halves[br].exchange = make_exchange(br)
    for st in [full] + halves:
        fill(st)
    for i in (0, 7, 20):
        for st in halves:  # restart every step from the batched state (no drift between the two runs)
            st.latents.copy_(full.latents)
        full.step(i)
        for st in halves:
            st.run_model(i)
        for st in halves:
            st.finish()
        torch.cuda.synchronize()
        assert torch.equal(halves[0].latents, halves[1].latents)
        if exact:
            assert torch.equal(halves[0].noise, full.noise), (order, i, rel(halves[0].noise, full.noise))
            assert torch.equal(halves[0].latents, full.latents)
        else:
            assert rel(halves[0].noise, full.noise) < 2e-2, (order, i, rel(halves[0].noise, full.noise))
            assert rel(halves[0].latents, full.latents) < 1e-2
    with pytest.raises(ValueError):
        pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=False, cfg_branch=0)


@pytest.mark.parametrize("order", ["s_major", "b_major"])
def test_two_clips_per_gpu_cfg_batch_4(order):
    """Two clips in one batch (CFG batch 4): the s-major time_context pairing then runs modulo 4
     and the per-clip temporal statistics must stay separate."""
    from ctrlv_b200 import pipeline
    from oracle import sampling as S
    from oracle import svd_oracle as O
    ou, oc, mu, mc = _build(order, dict(O.TINY_CONFIG))
    T, h, w, steps = 3, 8, 16, 25
    a = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], seed=1234, device=dev)
    b = S.make_inputs(T=T, h=h, w=w, xdim=O.TINY_CONFIG["cross_attention_dim"], seed=99, device=dev)
    cat2 = lambda k: torch.cat([a[k], b[k]])
    inp = dict(latents=cat2("latents"),
               image_latents=torch.cat([torch.zeros_like(cat2("image_latents_cond")), cat2("image_latents_cond")])
               .unsqueeze(1).repeat(1, T, 1, 1, 1),
               image_embeddings=torch.cat([torch.zeros_like(cat2("image_embeds_cond")), cat2("image_embeds_cond")]),
               cond_em=torch.cat([torch.zeros_like(cat2("cond_em_cond")), cat2("cond_em_cond")]),
               added_time_ids=torch.tensor([[6.0, 127.0, 0.02]], device=dev).repeat(4, 1), guidance=a["guidance"])
    trace = []
    with torch.no_grad():
        S.sample_loop(ou, oc, inp, num_steps=steps, trace=trace)
    pipe = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)
    out = pipe(cond_images=cat2("cond_em_cond"), height=h * 8, width=w * 8, num_frames=T, num_inference_steps=steps,
               latents=cat2("latents").clone(), output_type="latent", image_embeddings=cat2("image_embeds_cond"),
               image_latents=cat2("image_latents_cond"))
    assert tuple(out.frames.shape) == (2, T, 4, h, w)
    st = next(iter(pipe._steps.values()))
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(steps)
    prevs = [inp["latents"] * sch.init_noise_sigma] + trace[:-1]
    for i in range(0, steps, 3):
        st.latents.copy_(prevs[i]); st.step(i)
        assert rel(st.latents, trace[i]) < 1e-2, (order, i, rel(st.latents, trace[i]))
    # the two clips do not leak into each other: clip 0 alone gives the same result as clip 0 in the batch
    # (b-major order only: the 0.27.2 s-major pairing deliberately mixes contexts across the batch)
    if order == "b_major":
        solo = pipe(cond_images=a["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T, num_inference_steps=steps,
                    latents=a["latents"].clone(), output_type="latent", image_embeddings=a["image_embeds_cond"],
                    image_latents=a["image_latents_cond"])
        assert rel(solo.frames[0], out.frames[0]) < 2e-2


def test_fp16_and_fp32_inputs_like_autocast_callers(tiny):
    """callers may run the modules under torch.autocast(fp16) with fp32 or
    fp16 tensors; the drop-in fixes its compute dtype and casts at entry/exit (output dtype = input)."""
    _, _, mu, mc = tiny
    inp, x = _inputs(3, 8, 16, 2.0)
    t = torch.tensor(0.17, device=dev)
    outs = {}
    for dt in (torch.float32, torch.float16, torch.bfloat16):
        with torch.autocast("cuda", dtype=torch.float16, enabled=dt is torch.float16):
            d, m = mc(x.to(dt), timestep=t, encoder_hidden_states=inp["image_embeddings"].to(dt),
                      added_time_ids=inp["added_time_ids"].to(dt), control_cond=inp["cond_em"].to(dt), return_dict=False)
            y = mu(sample=x.to(dt), timestep=t, encoder_hidden_states=inp["image_embeddings"].to(dt),
                   added_time_ids=inp["added_time_ids"].to(dt), down_block_additional_residuals=d,
                   mid_block_additional_residuals=m, return_dict=False)[0]
        assert y.dtype == dt and m.dtype == dt and all(r.dtype == dt for r in d)
        outs[dt] = y.float()
    # (the inputs themselves are rounded to the caller's dtype: a few 1e-3 of input noise, amplified by the
    #  random-init tiny network)
    assert rel(outs[torch.float16], outs[torch.float32]) < 3e-2
    assert rel(outs[torch.bfloat16], outs[torch.float32]) < 5e-2


def test_groupnorm_statistics_from_epilogues_match_the_two_pass_path(tiny):
    """with models.GN_FUSED the GroupNorm statistics come from the epilogues of the launches
    that produce each norm's input; the result agrees with the round-1 path (one statistics pass per norm), is
    bit-reproducible, and launches fewer kernels."""
    from ctrlv_b200 import _lib, models
    ou, oc, mu, mc = tiny
    lib = _lib.load()
    inp, x = _inputs(4, 16, 16, 7.0)
    t = torch.tensor(0.486, device=dev)

    def run():
        n0 = lib.ctrlv_launch_count()
        md, mm = mc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        y = mu(x, t, inp["image_embeddings"], inp["added_time_ids"], md, mm, return_dict=False)[0]
        torch.cuda.synchronize()
        return y, md, lib.ctrlv_launch_count() - n0

    assert models.GN_FUSED
    y1, d1, n1 = run()
    y1b, _, _ = run()
    models.GN_FUSED = False
    try:
        y0, d0, n0 = run()
    finally:
        models.GN_FUSED = True
    assert torch.equal(y1, y1b)
    # the two paths round mean / rstd differently in the last bits; through ~150 bf16 layers that decorrelates the
    # rounding noise, so they agree to the bf16 noise level of the model (each is within 2e-2 of the fp32 oracle),
    # the ControlNet residuals (half the depth) more closely
    with torch.no_grad():
        od, om = oc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        oy = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], od, om, return_dict=False)[0]
    assert rel(y1, oy) < 2e-2 and rel(y0, oy) < 2e-2
    assert rel(y1, y0) < 2e-2 and max(rel(a, b) for a, b in zip(d1, d0)) < 2e-2
    assert n1 < n0 - 40, (n1, n0)  # the tiny widths fuse every norm of >= 128 channels


@pytest.mark.parametrize("L", [3, 77])
def test_multi_token_context_cross_attention(tiny, L):
    """controlnet.py:230,244-245: `encoder_hidden_states` is [batch, tokens, dim]; the pipelines pass one token
    (folded into a per-sample vector), a longer context runs real cross-attention (ctrlv_cross_attn) in every
    spatial and temporal transformer block, with the 0.27.2 S-major `time_context` order."""
    ou, oc, mu, mc = tiny
    inp, x = _inputs(3, 8, 16, 7.0)
    t = torch.tensor(0.486, device=dev)
    g = torch.Generator(device="cpu").manual_seed(5)
    ehs = torch.randn(2, L, inp["image_embeddings"].shape[-1], generator=g).to(dev)
    with torch.no_grad():
        od, om = oc(x, t, ehs, inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        oy = ou(x, t, ehs, inp["added_time_ids"], od, om, return_dict=False)[0]
        oy1 = ou(x, t, ehs[:, :1], inp["added_time_ids"], od, om, return_dict=False)[0]
    md, mm = mc(x, t, ehs, inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
    my = mu(x, t, ehs, inp["added_time_ids"], md, mm, return_dict=False)[0]
    torch.cuda.synchronize()
    assert max(rel(a, b) for a, b in zip(md, od)) < 2.5e-2 and rel(mm, om) < 2.5e-2
    assert rel(my, oy) < 2e-2
    assert rel(oy, oy1) > 1e-2  # the extra tokens matter in this test


def test_step_replayed_from_a_library_plan_equals_eager_and_graph(tiny):
    """ctrlv_plan_create / _finish / _run — the whole step (ControlNet on the side stream, UNet, CFG +
    Euler) recorded once and re-issued from C, arguments and tensor maps frozen at record time.  Bit-identical to
    the eager step and to the CUDA-graph replay, also after unrelated allocations recycle the caching allocator."""
    from ctrlv_b200 import _lib, pipeline
    from oracle import sampling as S
    ou, oc, mu, mc = tiny
    T, h, w = 4, 16, 16
    inp = S.make_inputs(T=T, h=h, w=w, xdim=mu.cfg["cross_attention_dim"], device=dev)
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(6)
    outs = {}
    for mode in ("eager", "graph", "plan"):
        st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=(mode == "graph"), use_plan=(mode == "plan"))
        st.set_schedule(sch.sigmas, sch.timesteps)
        st.image_latents.copy_(inp["image_latents"]); st.cond_em.copy_(inp["cond_em"])
        st.ehs.copy_(inp["image_embeddings"].reshape(2, -1)); st.added_time_ids.copy_(inp["added_time_ids"])
        st.guidance.copy_(inp["guidance"])
        st.latents.copy_(inp["latents"] * float(sch.init_noise_sigma))
        st.capture()
        if mode == "plan":
            n = _lib.load().ctrlv_plan_size(st._plan)
            assert n > 300, n  # every launch of the step is in the plan
        junk = [torch.randn(1 << 20, device=dev) for _ in range(6)]
        for i in range(6):
            st.step(i)
        torch.cuda.synchronize()
        outs[mode] = st.latents.clone()
        del junk
    assert torch.isfinite(outs["plan"]).all()
    assert torch.equal(outs["plan"], outs["eager"]) and torch.equal(outs["graph"], outs["eager"])
