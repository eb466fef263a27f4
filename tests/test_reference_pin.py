"""Pins the oracle's restatement of the reference's IN-REPO hot-path code against that code itself.

`tests/golden/ref_shim.py` loads /root/reference/src/ctrlv/{models,pipelines}/*.py as they lie and
runs them on a stand-in `diffusers` whose blocks are the oracle's (diffusers==0.27.2 itself is
absent), so every difference between the reference's forward / __call__ and the oracle's
`ControlNetModel.forward`, `UNetSpatioTemporalConditionModel.forward`, `sampling.sample_loop*`
shows up as a non-zero difference here.  Two layers:

* live tests (skipped where /root/reference is absent, e.g. on the GPU box): reference vs oracle, exact;
* golden tests (run everywhere): the oracle alone against `tests/golden/ref_forward.pt`, the
  vectors the reference code produced (`tests/golden/make_ref_golden.py`).
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import sampling as S  # noqa: E402
from oracle import svd_oracle as O  # noqa: E402
from tests.golden import make_ref_golden as G  # noqa: E402
from tests.golden import ref_shim as R  # noqa: E402

live = pytest.mark.skipif(not R.available(), reason="/root/reference is not present on this machine")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_forward.pt")


@pytest.fixture(scope="module")
def ref():
    r = R.load_reference()
    yield r
    R.unload()


@pytest.fixture(scope="module")
def models():
    return G.oracle_models()


oracle_pipeline_inputs = G.conditioning


# ------------------------------------------------------------------------------------------------
# live: the reference's code against the oracle
# ------------------------------------------------------------------------------------------------
@live
def test_reference_constructors_build_the_oracle_layout(ref, models):
    ou, oc = models
    rc = ref.ControlNetModel(**O.TINY_CONFIG)
    ru = ref.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG)
    for r, o in ((rc, oc), (ru, ou)):
        rs, os_ = r.state_dict(), o.state_dict()
        assert set(rs) == set(os_)
        assert all(rs[k].shape == os_[k].shape for k in rs)
    # controlnet.py:148-185: 1 + (2 + 1) * 3 + 2 zero-convs, all zero at construction
    assert len(rc.controlnet_down_blocks) == len(oc.controlnet_down_blocks) == 12
    assert all(float(m.weight.abs().max()) == 0.0 and float(m.bias.abs().max()) == 0.0
               for m in list(rc.controlnet_down_blocks) + [rc.controlnet_mid_block])
    assert [tuple(m.weight.shape) for m in rc.controlnet_down_blocks] == \
           [tuple(m.weight.shape) for m in oc.controlnet_down_blocks]
    # the config the reference records (register_to_config) carries what the oracle's does
    for k in ("in_channels", "block_out_channels", "num_attention_heads", "cross_attention_dim", "num_frames",
              "addition_time_embed_dim", "projection_class_embeddings_input_dim", "layers_per_block"):
        assert getattr(rc.config, k) == getattr(oc.config, k), k


@live
@pytest.mark.parametrize("timestep", [torch.tensor(1.25), torch.tensor([1.25, 1.25])], ids=["t0d", "tB"])
@pytest.mark.parametrize("scale", [1.0, 0.7])
def test_reference_forwards_equal_oracle_forwards(ref, models, timestep, scale):
    ou, oc = models
    ru, rc = G.reference_models(ref, ou, oc)
    fi = G.forward_inputs()
    args = (fi["sample"], timestep, fi["encoder_hidden_states"], fi["added_time_ids"])
    with torch.no_grad():
        rd, rm = rc(*args, control_cond=fi["control_cond"], conditioning_scale=scale, return_dict=False)
        od, om = oc(*args, control_cond=fi["control_cond"], conditioning_scale=scale, return_dict=False)
        assert len(rd) == len(od) == 12
        for a, b in zip(rd, od):
            assert torch.equal(a, b)
        assert torch.equal(rm, om)
        ro = rc(*args, control_cond=fi["control_cond"], conditioning_scale=scale)  # return_dict=True
        assert torch.equal(ro.mid_block_res_sample, om) and torch.equal(ro.down_block_res_samples[5], od[5])
        r1 = ru(*args, down_block_additional_residuals=rd, mid_block_additional_residuals=rm, return_dict=False)[0]
        o1 = ou(*args, down_block_additional_residuals=od, mid_block_additional_residuals=om, return_dict=False)[0]
        assert r1.shape == (2, G.T, 4, G.H, G.W) and torch.equal(r1, o1)
        assert torch.equal(ru(*args).sample, ou(*args).sample)  # residuals None: plain SVD UNet (:61)
        # only one of the two given: the reference ignores both (is_controlnet needs both, :61)
        assert torch.equal(ru(*args, down_block_additional_residuals=rd).sample, ou(*args).sample)
        assert torch.equal(ou(*args, down_block_additional_residuals=od).sample, ou(*args).sample)


@live
def test_reference_from_unet_equals_oracle_from_unet(ref, models):
    ou, oc = models
    ru, _ = G.reference_models(ref, ou, oc)
    rc2, oc2 = ref.ControlNetModel.from_unet(ru), O.ControlNetModel.from_unet(ou)
    rs, os_ = rc2.state_dict(), oc2.state_dict()
    assert set(rs) == set(os_)
    shared = [k for k in rs if k in ou.state_dict()]
    assert len(shared) > 100 and not any(k.startswith("control") for k in shared)
    for k in shared:
        assert torch.equal(rs[k], os_[k]) and torch.equal(rs[k], ou.state_dict()[k])
    for k in rs:
        if k.startswith("controlnet_"):
            assert float(rs[k].abs().max()) == 0.0 and float(os_[k].abs().max()) == 0.0


@live
@pytest.mark.parametrize("cond_channels,scale,batch", [(4, 1.0, 1), (3, 0.5, 1), (4, 1.0, 2)],
                         ids=["latent-cond", "frame-cond", "two-clips"])
def test_reference_control_pipeline_equals_oracle_loop(ref, models, cond_channels, scale, batch):
    ou, oc = models
    ru, rc = G.reference_models(ref, ou, oc)
    image, cond, latents = G.pipeline_inputs(batch=batch, cond_channels=cond_channels, seed=20 + cond_channels)
    with torch.no_grad():
        got = G.run_control_pipeline(ref, ru, rc, image, cond, latents, scale=scale)
        want = S.sample_loop(ou, oc, oracle_pipeline_inputs(image, cond, latents), num_steps=G.STEPS,
                             conditioning_scale=scale)
    assert got.shape == latents.shape and torch.isfinite(got).all()
    assert torch.equal(got, want)


@live
@pytest.mark.parametrize("ncond", [1, 3])
def test_reference_bbox_predictor_pipeline_equals_oracle_loop(ref, models, ncond):
    ou, _ = models
    ru, _ = G.reference_models(ref, ou, models[1])
    image, condf, latents = G.pipeline_inputs(cond_channels=3, seed=31)
    with torch.no_grad():
        got = G.run_bbox_pipeline(ref, ru, image, condf, latents, num_cond_bbox_frames=ncond)
        want = S.sample_loop_bbox_predictor(ou, oracle_pipeline_inputs(image, condf, latents), num_steps=G.STEPS,
                                            num_cond_bbox_frames=ncond)
    assert torch.equal(got, want)


@live
def test_reference_pipeline_error_conventions(ref, models):
    ou, oc = models
    ru, rc = G.reference_models(ref, ou, oc)
    image, cond, latents = G.pipeline_inputs()
    pipe = ref.StableVideoControlPipeline(vae=R.FakeVAE(), image_encoder=R.FakeImageEncoder(G.XDIM), unet=ru,
                                          controlnet=rc, scheduler=R.SchedulerShim(), feature_extractor=None)
    with pytest.raises(ValueError):  # pipeline_video_control.py:66-67
        pipe(image, cond_images=cond, height=60, width=64, num_frames=G.T, latents=latents, output_type="latent")
    with pytest.raises(ValueError):  # :59-63
        pipe(image, cond_images=[1, 2], height=64, width=64, num_frames=G.T, latents=latents, output_type="latent")
    with pytest.raises(AssertionError):  # :87
        pipe(image, cond_images=torch.zeros(1, G.T, 5, G.H, G.W), height=64, width=64, num_frames=G.T,
             latents=latents, output_type="latent")
    with pytest.raises(ValueError):  # controlnet.py:80-83
        ref.ControlNetModel(block_out_channels=(64, 128))
    with pytest.raises(ValueError):
        O.ControlNetModel(block_out_channels=(64, 128))


# ------------------------------------------------------------------------------------------------
# golden: the oracle alone against what the reference code produced (runs on any machine)
# ------------------------------------------------------------------------------------------------
def _close(a, b):
    return torch.allclose(a, b, rtol=1e-4, atol=1e-5)  # fp32 CPU kernels may differ across hosts


def test_oracle_reproduces_reference_forward_vectors(models):
    gold = torch.load(GOLDEN)
    ou, oc = models
    fi = G.forward_inputs()
    args = (fi["sample"], fi["timestep"], fi["encoder_hidden_states"], fi["added_time_ids"])
    with torch.no_grad():
        down, mid = oc(*args, control_cond=fi["control_cond"], conditioning_scale=0.7, return_dict=False)
        stats = torch.stack([torch.stack([d.norm(), d.sum()]) for d in down])
        assert torch.allclose(stats, gold["ctrl_down_stats"], rtol=1e-3, atol=1e-4)
        assert _close(down[3], gold["ctrl_down_3"]) and _close(down[11], gold["ctrl_down_11"])
        assert _close(mid, gold["ctrl_mid"])
        out = ou(*args, down_block_additional_residuals=down, mid_block_additional_residuals=mid, return_dict=False)[0]
        assert _close(out, gold["unet_with_residuals"])
        assert _close(ou(*args).sample, gold["unet_plain"])


def test_oracle_reproduces_reference_pipeline_vectors(models):
    gold = torch.load(GOLDEN)
    ou, oc = models
    rel = lambda a, b: float((a - b).norm() / b.norm())
    with torch.no_grad():
        image, cond, latents = G.pipeline_inputs()
        got = S.sample_loop(ou, oc, oracle_pipeline_inputs(image, cond, latents), num_steps=G.STEPS)
        assert rel(got, gold["control_pipeline_latents"]) < 1e-4
        image, condf, latents = G.pipeline_inputs(cond_channels=3, seed=13)
        inp = oracle_pipeline_inputs(image, condf, latents)
        got = S.sample_loop(ou, oc, inp, num_steps=G.STEPS, conditioning_scale=0.5)
        assert rel(got, gold["control_pipeline_latents_from_frames"]) < 1e-4
        got = S.sample_loop_bbox_predictor(ou, inp, num_steps=G.STEPS, num_cond_bbox_frames=1)
        assert rel(got, gold["bbox_pipeline_latents"]) < 1e-4


# ------------------------------------------------------------------------------------------------
# row f-3: the image-embedding prologue against the reference's own `encode_video_image`
# ------------------------------------------------------------------------------------------------
def _reference_functions(relpath, names, extra_ns=None):
    """Function definitions taken from a reference source file's AST and executed as they are (the
    modules import diffusers at the top and cannot be imported whole)."""
    import ast
    path = os.path.join(R.REFERENCE_SRC, relpath)
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    ns = {"torch": torch}
    ns.update(extra_ns or {})
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    assert set(names) <= set(ns)
    return ns


@live
def test_reference_encode_video_image_equals_oracle_encode_image():
    """src/ctrlv/utils/util.py:97-125 (antialiased resize to 224, un-normalise + clamp, CLIP
    normalisation by the real `CLIPImageProcessor`, `image_encoder(...).image_embeds`) with a small
    random `transformers` CLIP tower, against `clip_oracle.encode_image(clamp=True)`."""
    tf = pytest.importorskip("transformers")
    from oracle import clip_oracle as CO
    resize = _reference_functions("bbox_generator_baseline/utils/image_encoder.py",
                                  {"_resize_with_antialiasing", "_compute_padding", "_filter2d", "_gaussian",
                                   "_gaussian_blur2d"})
    ref = _reference_functions("utils/util.py", {"encode_video_image"},
                               {"_resize_with_antialiasing": resize["_resize_with_antialiasing"]})
    cfg = dict(CO.TINY_CLIP_CONFIG, image_size=224, patch_size=32)
    torch.manual_seed(3)
    hf = tf.CLIPVisionModelWithProjection(tf.CLIPVisionConfig(**cfg)).eval()
    mine = CO.CLIPVisionModelWithProjection(**cfg).eval()
    mine.load_state_dict({k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")})
    g = torch.Generator().manual_seed(4)
    for shape in ((2, 3, 320, 512), (1, 3, 96, 160)):
        image01 = torch.rand(shape, generator=g)
        with torch.no_grad():
            want = ref["encode_video_image"](image01 * 2.0 - 1.0, tf.CLIPImageProcessor(), torch.float32, hf)
            got = CO.encode_image(mine, image01, clamp=True)
        assert got.shape == (shape[0], 1, cfg["projection_dim"]) and want.shape == (shape[0], cfg["projection_dim"])
        assert torch.allclose(got[:, 0], want, atol=2e-5, rtol=1e-4), float((got[:, 0] - want).abs().max())


@live
def test_product_host_helpers_equal_reference_helpers(ref):
    """Host-side glue of this repo's pipeline (no CUDA involved) against the reference's own helpers:
    `get_add_time_ids` (src/ctrlv/utils/util.py:147-170), `_encode_vae_condition`'s layout rules and
    `check_inputs` (pipeline_video_control.py:51-68)."""
    from types import SimpleNamespace
    from ctrlv_b200.pipeline import StableVideoControlPipeline as Mine
    util = _reference_functions("utils/util.py", {"get_add_time_ids", "get_model_attr"}, {"nn": torch.nn})
    unet = SimpleNamespace(config=SimpleNamespace(addition_time_embed_dim=256),
                           add_embedding=SimpleNamespace(linear_1=SimpleNamespace(in_features=768)))
    me = Mine.__new__(Mine)
    me.unet = unet
    for fps, mb, na, bs in ((6, 127, 0.02, 1), (24, 40, 0.1, 3)):
        want = util["get_add_time_ids"](fps, mb, na, torch.float32, bs, unet)
        got = me._get_add_time_ids(fps, mb, na, bs)
        assert got.dtype == want.dtype and torch.equal(got, want)
    unet.add_embedding.linear_1.in_features = 512  # mis-configured model: same exception type
    with pytest.raises(ValueError):
        util["get_add_time_ids"](6, 127, 0.02, torch.float32, 1, unet)
    with pytest.raises(ValueError):
        me._get_add_time_ids(6, 127, 0.02, 1)
    # check_inputs: the reference's own method (through the shim) and ours reject the same calls
    theirs = ref.StableVideoControlPipeline.__new__(ref.StableVideoControlPipeline)
    img, cond = torch.zeros(1, 3, 64, 64), torch.zeros(1, 2, 4, 8, 8)
    for args in ((img, cond, 64, 64), (img, cond, 60, 64), (img, cond, 64, 36), (3.0, cond, 64, 64),
                 (img, [cond], 64, 64), (img, None, 64, 64)):
        outcome = []
        for p in (theirs, me):
            try:
                p.check_inputs(*args)
                outcome.append("ok")
            except ValueError:
                outcome.append("ValueError")
        assert outcome[0] == outcome[1], (args[2:], type(args[0]), type(args[1]), outcome)
