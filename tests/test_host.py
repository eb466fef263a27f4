"""CPU tests of the host-side logic and of the C-ABI library surface (no compute calls)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_spec_matches_oracle_state_dict():
    from ctrlv_b200 import models
    from oracle import svd_oracle as O
    for over in ({}, O.TINY_CONFIG):
        with torch.device("meta"):
            u = O.UNetSpatioTemporalConditionModel(**over)
            c = O.ControlNetModel(**over)
        for mod, ctrl in ((u, False), (c, True)):
            spec = models.param_spec({**models.SVD_CONFIG, **over}, ctrl)
            sd = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
            assert set(spec) == set(sd)
            assert all(tuple(spec[k]) == sd[k] for k in sd)
    import math
    assert sum(math.prod(s) for s in models.param_spec(models.SVD_CONFIG, False).values()) == 1_524_623_082
    assert sum(math.prod(s) for s in models.param_spec(models.SVD_CONFIG, True).values()) == 680_946_897


def test_product_scheduler_equals_oracle_scheduler():
    from ctrlv_b200.pipeline import EulerDiscreteScheduler
    from oracle.sampling import EulerDiscreteSchedulerOracle
    for n in (2, 25, 30, 50):
        a = EulerDiscreteScheduler().set_timesteps(n)
        b = EulerDiscreteSchedulerOracle()
        b.set_timesteps(n)
        assert torch.equal(a.sigmas, b.sigmas)
        assert torch.allclose(a.timesteps, b.timesteps, rtol=0, atol=2e-7)
        assert abs(a.init_noise_sigma - float(b.init_noise_sigma)) < 1e-3


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ctrlv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctrlv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ctrlv_b200 import _lib
    names = _declared_functions()
    assert len(names) >= 20
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    lib = _lib.load()  # builds in-tree if absent (nvcc cross-compiles without a GPU)
    for n in names:
        assert hasattr(lib, n), n
    assert b"sm_100a" in lib.ctrlv_version()
    assert isinstance(lib.ctrlv_last_error(), bytes)


def test_struct_layout_matches_header():
    """ctypes mirrors of the C structs: sizes must match what nvcc compiled (checked through a tiny
    C program compiled with gcc against the same header)."""
    import subprocess
    import tempfile
    from ctrlv_b200 import _lib
    prog = r'''
#include <stdio.h>
#include "ctrlv_b200.h"
int main(void){ printf("%zu %zu %zu %zu\n", sizeof(ctrlv_epilogue), sizeof(ctrlv_src), sizeof(ctrlv_seg), sizeof(ctrlv_igemm_desc)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(_lib.Epilogue), ctypes.sizeof(_lib.Src), ctypes.sizeof(_lib.Seg),
                     ctypes.sizeof(_lib.IgemmDesc)]


def test_input_validation_mirrors_reference():
    from ctrlv_b200 import models
    chk = models._PackedModel._check_inputs
    ok = torch.zeros(2, 3, 8, 16, 16)
    ehs = torch.zeros(2, 1, 1024)
    ids = torch.zeros(2, 3)
    chk(ok, torch.tensor(1.0), ehs, ids)
    with pytest.raises(TypeError):
        chk(ok, 1.0, ehs, ids)  # controlnet.py:262-264: timestep must be a tensor
    with pytest.raises(ValueError):
        chk(torch.zeros(2, 3, 8, 12, 16), torch.tensor(1.0), ehs, ids)
    chk(ok, torch.tensor(1.0), torch.zeros(2, 4, 1024), ids)  # multi-token contexts are accepted (ctrlv_cross_attn)
    with pytest.raises(ValueError):
        chk(ok, torch.tensor(1.0), torch.zeros(2, 1024), ids)
    with pytest.raises(NotImplementedError):
        chk(ok, torch.tensor(1.0), torch.zeros(2, 300, 1024), ids)


def test_pipeline_check_inputs():
    from ctrlv_b200.pipeline import StableVideoControlPipeline
    p = StableVideoControlPipeline.__new__(StableVideoControlPipeline)
    with pytest.raises(ValueError):  # pipeline_video_control.py:61-65
        p.check_inputs(None, None, 320, 512)
    with pytest.raises(ValueError):  # :67-68
        p.check_inputs(None, torch.zeros(1), 321, 512)
    p.check_inputs(None, torch.zeros(1), 320, 512)


def test_shard_clips():
    from ctrlv_b200.parallel import shard_clips
    for n in (1, 7, 8, 13):
        for w in (1, 2, 4, 8):
            parts = [shard_clips(n, w, r) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_clips(4, 2, 2)


def test_geglu_interleave_and_conv_packing():
    from ctrlv_b200 import models
    w = torch.arange(8 * 3, dtype=torch.float32).reshape(8, 3)
    b = torch.arange(8, dtype=torch.float32)
    wi, bi = models._interleave_geglu(w, b)
    assert torch.equal(wi[0::2], w[:4]) and torch.equal(wi[1::2], w[4:])
    assert torch.equal(bi[0::2], b[:4]) and torch.equal(bi[1::2], b[4:])
    cw = torch.randn(5, 7, 3, 3)
    p = models._conv9(cw)
    assert p.shape == (5, 63) and torch.equal(p[:, 7 * 4:7 * 5], cw[:, :, 1, 1])  # tap (1,1) = centre
    tw = torch.randn(5, 7, 3, 1, 1)
    assert torch.equal(models._conv_t3(tw)[:, 7:14], tw[:, :, 1, 0, 0])


def test_safetensors_reader_writer_roundtrip_and_interop(tmp_path):
    """f-4: the self-contained safetensors codec against the `safetensors` package (both ways)."""
    import torch
    from ctrlv_b200 import checkpoint
    g = torch.Generator().manual_seed(0)
    sd = {"a.weight": torch.randn(3, 5, generator=g), "b.bias": torch.randn(7, generator=g).to(torch.bfloat16),
          "c": torch.randn(2, 2, 2, generator=g).to(torch.float16), "n": torch.arange(6).reshape(2, 3),
          "scalar": torch.tensor(1.5), "empty": torch.zeros(0, 4)}
    p = str(tmp_path / "x.safetensors")
    checkpoint.write_safetensors(p, sd, {"format": "pt"})
    back = checkpoint.read_safetensors(p)
    assert set(back) == set(sd)
    for k in sd:
        assert back[k].dtype == sd[k].dtype and torch.equal(back[k], sd[k]), k
    st = pytest.importorskip("safetensors.torch")
    lib = st.load_file(p)
    for k in sd:
        assert torch.equal(lib[k], sd[k]), k
    p2 = str(tmp_path / "y.safetensors")
    st.save_file({k: v.contiguous() for k, v in sd.items()}, p2)
    back2 = checkpoint.read_safetensors(p2)
    for k in sd:
        assert torch.equal(back2[k], sd[k]), k
    with open(str(tmp_path / "bad.safetensors"), "wb") as f:
        f.write(b"\xff" * 16)
    with pytest.raises(ValueError):
        checkpoint.read_safetensors(str(tmp_path / "bad.safetensors"))


@pytest.mark.parametrize("mode", ["single", "sharded", "bin", "variant"])
def test_diffusers_dir_layouts(tmp_path, mode):
    import torch
    from ctrlv_b200 import checkpoint
    sd = {f"down_blocks.{i}.w": torch.full((4, 4), float(i)) for i in range(5)}
    cfg = {"in_channels": 8, "block_out_channels": (64, 128)}
    kw = dict(single={}, sharded=dict(max_shard_bytes=100), bin=dict(safe_serialization=False),
              variant=dict(variant="fp16"))[mode]
    checkpoint.save_diffusers_dir(str(tmp_path), cfg, sd, "ControlNetModel", subfolder="controlnet", **kw)
    names = sorted(os.listdir(str(tmp_path / "controlnet")))
    assert "config.json" in names
    if mode == "sharded":
        assert any(n.endswith(".index.json") for n in names) and sum(n.endswith(".safetensors") for n in names) > 1
    config, back = checkpoint.load_diffusers_dir(str(tmp_path), "controlnet", kw.get("variant"))
    assert config["_class_name"] == "ControlNetModel" and config["block_out_channels"] == [64, 128]
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    with pytest.raises(OSError):
        checkpoint.load_diffusers_dir(str(tmp_path), "unet")


def test_subpixel_upsample_conv_algebra_cpu():
    """conv3x3(nearest2x(x)) == four 2x2 phase convs of x with the summed taps of ops.pack_upconv3x3
    (pure fp32 on the CPU: pins the weight packing and the patch/offset convention of
    ctrlv_upsample2x_conv3x3 without a GPU)."""
    import torch.nn.functional as F
    from ctrlv_b200 import ops
    g = torch.Generator().manual_seed(0)
    N, Cc, H, W = 5, 3, 6, 7
    x = torch.randn(2, Cc, H, W, generator=g)
    w = torch.randn(N, Cc, 3, 3, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, padding=1)
    wp = ops.pack_upconv3x3(w, device="cpu", dtype=torch.float32)          # [4, N, 4*C]
    assert tuple(wp.shape) == (4, N, 4 * Cc)
    xp = F.pad(x, (1, 1, 1, 1))                                            # zero halo = TMA out-of-bounds fill
    got = torch.zeros_like(want)
    for py in (0, 1):
        for px in (0, 1):
            acc = torch.zeros(2, N, H, W)
            i = 0
            for dy in ((-1, 0) if py == 0 else (0, 1)):
                for dx in ((-1, 0) if px == 0 else (0, 1)):
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]  # x(y + dy, x + dx)
                    acc += torch.einsum("nc,bchw->bnhw", wp[py * 2 + px][:, i * Cc:(i + 1) * Cc], patch)
                    i += 1
            got[:, :, py::2, px::2] = acc
    assert torch.allclose(got, want, atol=1e-5), float((got - want).abs().max())


def test_device_polynomial_constants():
    """The two fitted approximations used on the device, checked on the CPU with the constants read
    from the sources: the tanh-form erf-GELU of the GEGLU epilogue (max abs error 2.5e-5) and the cubic
    2^f of the attention softmax's FMA-pipe exponentials (max rel. error 7.5e-5 on [-0.5, 0.5])."""
    import math
    import numpy as np
    src = open(os.path.join(ROOT, "ctrl-v_b200", "csrc", "common.cuh")).read()
    body = src[src.index("void geglu2_f("):]
    c2, c1 = (float(v) for v in re.findall(r"f2_splat\((-?[0-9.eE+-]+)f\)", body[:body.index("f2_fma(u, p")])[:2])
    c0 = float(re.search(r"p = f2_fma\(u, p, f2_splat\((-?[0-9.eE+-]+)f\)\)", body).group(1))
    g = np.linspace(-12.0, 12.0, 480001)
    u = np.minimum(g * g, 64.0)
    approx = 0.5 * g * (1.0 + np.tanh(g * (c0 + u * (c1 + u * c2))))
    exact = 0.5 * g * (1.0 + np.vectorize(math.erf)(g / math.sqrt(2.0)))
    assert float(np.abs(approx - exact).max()) < 3e-5
    fl = open(os.path.join(ROOT, "ctrl-v_b200", "csrc", "flash.cu")).read()
    seg = fl[fl.index("uint64_t q = f2_fma(ff"):]
    k3, k2 = (float(v) for v in re.findall(r"f2_splat\((-?[0-9.eE+-]+)f\)", seg[:seg.index(";")]))
    k1 = float(re.search(r"q = f2_fma\(ff, q, f2_splat\((0\.69[0-9]+)f\)\)", seg).group(1))
    k0 = float(re.search(r"q = f2_fma\(ff, q, f2_splat\((0\.99[0-9]+)f\)\)", seg).group(1))
    f = np.linspace(-0.5, 0.5, 100001)
    p = ((k3 * f + k2) * f + k1) * f + k0
    assert float(np.abs(p / 2.0 ** f - 1.0).max()) < 1e-4
    # the exponent insert: 2^x = 2^f * 2^n with n = round(x) taken from the low mantissa bits of x + 1.5 * 2^23
    x = np.float32(-37.3)
    t = np.float32(x + np.float32(12582912.0))
    n = int(t.view(np.uint32)) - 0x4B400000
    assert n == round(float(x)) and abs(float(x) - n) <= 0.5


def test_fastdiv_header_exact_on_cpu(tmp_path):
    """ctrl-v_b200/csrc/fastdiv.h (tile-index decomposition of the implicit GEMM) against the hardware
    division: every divisor up to 4096 over a dense range of dividends plus the 32-bit corner values."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    src = tmp_path / "fd.cpp"
    src.write_text(r'''
#include <cstdio>
#include "fastdiv.h"
int main() {
  long bad = 0;
  const uint32_t corner[] = {0u, 1u, 2u, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu, 123456789u, 4000000000u};
  for (uint32_t d = 1; d <= 4096; ++d) {
    const ctrlv::FastDiv f = ctrlv::make_fastdiv(d);
    for (uint32_t n = 0; n < 300000; n += (d < 64 ? 1 : 7)) {
      uint32_t q, r;
      ctrlv::fd_divmod(n, f, q, r);
      if (q != n / d || r != n % d) ++bad;
    }
    for (uint32_t n : corner) if (ctrlv::fd_div(n, f) != n / d) ++bad;
  }
  std::printf("bad=%ld\n", bad);
  return bad != 0;
}
''')
    exe = tmp_path / "fd"
    inc = os.path.join(ROOT, "ctrl-v_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", inc, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout + out.stderr


def test_tile_plan_heuristics_cpu():
    """ctrlv_igemm_plan (no CUDA calls): the n-tile / cta_group choices for the step's characteristic
    problems on a 148-SM device, as measured best in profiles/r01_igemm_tile_sweep.json and, after the MMA warp
    started issuing several k-blocks per round, profiles/r02_igemm_tile_sweep.json."""
    from ctrlv_b200 import _lib
    lib = _lib.load(build_if_missing=False)

    def plan(X, Y, Z, K, N):
        d = _lib.IgemmDesc()
        d.nsrc = 1
        d.X, d.Y, d.Z, d.N, d.K = X, Y, Z, N, K
        d.nseg = 1
        d.seg[0].nchunk = K // 64
        box = (ctypes.c_int32 * 3)(); bn = ctypes.c_int32(); cg = ctypes.c_int32()
        assert lib.ctrlv_igemm_plan(ctypes.byref(d), 148, box, ctypes.byref(bn), ctypes.byref(cg)) == 0
        return tuple(box), bn.value, cg.value

    # level-0 GEGLU up-projection: widest tile, single CTAs (K = 320 loses with CTA pairs)
    assert plan(71680, 1, 1, 320, 2560)[1:] == (256, 1)
    # level-2 GEGLU (was BN=128 before the cost-model refit: 115 -> 86 us), level-1 QKV, level-1 residual linear
    assert plan(4480, 1, 1, 1280, 10240)[1:] == (256, 2)
    assert plan(17920, 1, 1, 640, 1920)[1:] == (192, 2)   # r02 sweep: pairs 48-50 us, single CTAs 52-54 us
    assert plan(17920, 1, 1, 640, 640)[1:] == (160, 1)
    # 3x3 convs: level 0 (K = 2880) pairs; level 3 (9 m-tiles, K = 11520) pairs since the refit
    box, bn, cg = plan(64, 40, 28, 2880, 320)
    assert box[0] * box[1] * box[2] <= 128 and (bn, cg) == (160, 2)
    assert plan(8, 5, 28, 11520, 1280)[2] == 2
    # level 3 (10 m-tiles): pairs for the long-K problems since the r02 refit (conv(3,1,1) 31.6 -> 27.7 us,
    # Linear 5120 -> 1280 35.8 -> 29.7 us); tiny M with short K and narrow N stays on single CTAs
    assert plan(40, 14, 2, 3840, 1280)[2] == 2
    assert plan(1120, 1, 1, 5120, 1280)[2] == 2
    assert plan(1120, 1, 1, 1280, 1280)[2] == 1
    # the box never exceeds one 128-row MMA tile and covers the row space
    for X, Y, Z in ((64, 40, 28), (16, 10, 28), (2560, 14, 2), (7, 3, 5)):
        b = plan(X, Y, Z, 64, 64)[0]
        assert 1 <= b[0] * b[1] * b[2] <= 128 and b[0] <= max(X, 128) and b[1] <= Y and b[2] <= Z


def test_scheduler_from_pretrained_config(tmp_path):
    """A Stable-Video-Diffusion directory's scheduler/scheduler_config.json is honoured; modes the
    fused CFG + Euler kernel does not implement are refused, not silently replaced."""
    import json
    from ctrlv_b200.pipeline import EulerDiscreteScheduler
    svd = {"_class_name": "EulerDiscreteScheduler", "_diffusers_version": "0.24.0.dev0", "beta_end": 0.012,
           "beta_schedule": "scaled_linear", "beta_start": 0.00085, "clip_sample": False,
           "interpolation_type": "linear", "num_train_timesteps": 1000, "prediction_type": "v_prediction",
           "set_alpha_to_one": False, "sigma_max": 700.0, "sigma_min": 0.002, "skip_prk_steps": True,
           "steps_offset": 1, "timestep_spacing": "leading", "timestep_type": "continuous",
           "trained_betas": None, "use_karras_sigmas": True}
    d = tmp_path / "svd" / "scheduler"
    d.mkdir(parents=True)
    (d / "scheduler_config.json").write_text(json.dumps(svd))
    a = EulerDiscreteScheduler.from_pretrained(str(tmp_path / "svd"), subfolder="scheduler").set_timesteps(25)
    b = EulerDiscreteScheduler().set_timesteps(25)
    assert torch.equal(a.sigmas, b.sigmas) and torch.equal(a.timesteps, b.timesteps)
    c = EulerDiscreteScheduler.from_config(dict(svd, sigma_max=80.0, timestep_spacing="trailing")).set_timesteps(10)
    assert abs(float(c.sigmas[0]) - 80.0) < 1e-4 and c.init_noise_sigma == pytest.approx(80.0, rel=1e-6)
    for bad in ({"prediction_type": "epsilon"}, {"use_karras_sigmas": False}, {"timestep_type": "discrete"},
                {"_class_name": "DDIMScheduler"}):
        with pytest.raises(NotImplementedError):
            EulerDiscreteScheduler.from_config(dict(svd, **bad))
    with pytest.raises(OSError):
        EulerDiscreteScheduler.from_pretrained(str(tmp_path / "nowhere"))


def test_conditioning_image_is_resized_like_the_vae_image_processor():
    """pipeline_video_control.py:227 `image_processor.preprocess(image, height, width)`: diffusers 0.27.2 resizes
    tensors with F.interpolate's default nearest rule and PIL images with PIL's Lanczos filter."""
    import numpy as np
    import torch.nn.functional as F
    from PIL import Image
    from ctrlv_b200.pipeline import StableVideoControlPipeline as P
    g = torch.Generator().manual_seed(0)
    for (H, W, h, w) in ((48, 80, 32, 64), (30, 50, 64, 96), (64, 64, 64, 64), (37, 53, 40, 24)):
        x = torch.rand(2, 3, H, W, generator=g)
        assert torch.equal(P._preprocess_image(x, h, w), F.interpolate(x, size=(h, w)))
    im = Image.fromarray((np.random.RandomState(0).rand(48, 80, 3) * 255).astype("uint8"))
    got = P._preprocess_image([im, im], 32, 64)
    want = torch.from_numpy(np.array(im.resize((64, 32), resample=Image.LANCZOS)).astype("float32") / 255.0).permute(2, 0, 1)
    assert got.shape == (2, 3, 32, 64) and torch.equal(got[0], want) and torch.equal(got[1], want)
    assert torch.equal(P._preprocess_image(im, 48, 80)[0],
                       torch.from_numpy(np.array(im).astype("float32") / 255.0).permute(2, 0, 1))


def test_groupnorm_statistics_table_geometry():
    """ops.GNStats: replication factor (power of two, >= 128 table rows in total) and which norm widths the
    producers' epilogues can serve (an aligned 8-column piece must fall into at most two groups; the residual-add
    kernel reduces channel pairs)."""
    from ctrlv_b200 import ops
    for units in (1, 2, 3, 28, 50, 127, 128, 129, 500):
        r = ops.GNStats.replicas(units)
        assert r & (r - 1) == 0 and r * units >= 128 and (r == 1 or (r // 2) * units < 128)
        assert ops.GNStats.numel(units) == r * units * 64
    for C, ok in ((64, False), (128, True), (192, True), (256, True), (320, True), (640, True), (960, True), (1280, True),
                  (1920, True), (2560, True), (160, False), (96, False)):
        assert ops.GNStats.fusable(C) == ok, C
    for C in (128, 192, 320, 960, 1920):  # every aligned 8-column piece of a fusable width spans <= 2 groups
        cg = C // 32
        for c0 in range(0, C, 8):
            assert len({(c0 + j) // cg for j in range(8)}) <= 2, (C, c0)


def test_new_entry_points_reject_bad_arguments_before_touching_the_device():
    """Argument validation of the round-2 entry points runs before any CUDA call: usable on a host without a GPU.
    (ctrlv_linear_ln / ctrlv_feedforward_ln / ctrlv_igemm_streamk; messages via ctrlv_last_error.)"""
    import ctypes as C
    from ctrlv_b200 import _lib
    lib = _lib.load()
    ep = _lib.Epilogue()
    ep.s_acc = 1.0
    ep.out, ep.ld_out = 0x1000, 960
    x, w = 0x2000, 0x3000  # never dereferenced: every call below fails its argument checks
    # N must be a multiple of 64, eps positive
    assert lib.ctrlv_linear_ln(x, 320, 128, 320, 1e-5, None, 0, 1, 1, w, 100, C.byref(ep), None) == -1
    assert b"multiple of 64" in lib.ctrlv_last_error()
    assert lib.ctrlv_linear_ln(x, 320, 128, 320, 0.0, None, 0, 1, 1, w, 960, C.byref(ep), None) == -1
    # the projection epilogue takes a bias and a bf16 out only
    ep.res1, ep.ld_res1 = 0x4000, 960
    assert lib.ctrlv_linear_ln(x, 320, 128, 320, 1e-5, None, 0, 1, 1, w, 960, C.byref(ep), None) == -1
    assert b"bias and a bf16 out only" in lib.ctrlv_last_error()
    ep.res1, ep.ld_res1 = None, 0
    # row-bias table in front of the norm: ld >= K, 16-byte aligned
    assert lib.ctrlv_linear_ln(x, 320, 128, 320, 1e-5, 0x5000, 64, 1, 1, w, 960, C.byref(ep), None) == -1
    assert lib.ctrlv_feedforward_ln(x, 320, 128, 320, 1e-5, 0x5004, 320, 1, 1, w, 0x6000, w, C.byref(ep), None) == -1
    assert b"row-bias" in lib.ctrlv_last_error()
    assert lib.ctrlv_feedforward_ln(x, 320, 128, 320, -1.0, None, 0, 1, 1, w, 0x6000, w, C.byref(ep), None) == -1
    # widths above 320 do not fit the 512 TMEM columns
    assert lib.ctrlv_linear_ln(x, 640, 128, 640, 1e-5, None, 0, 1, 1, w, 1920, C.byref(ep), None) == -1
    assert b"320" in lib.ctrlv_last_error()
    # stream-K hook: three modes
    assert lib.ctrlv_igemm_streamk(3) == -1
    assert lib.ctrlv_igemm_streamk(0) == 0


def test_streamk_schedule_arithmetic():
    """The index arithmetic of the stream-K schedule (csrc/igemm.cu: WorkIter<true>, the partial-slot rule of the
    GEMM kernel and the contributor range / slot parity igemm_fixup_kernel recomputes), restated in Python and
    checked over random problems: every k-block of every tile is covered exactly once, a tile is either whole
    (one CTA, no slot) or split into contiguous partial ranges whose CTAs are exactly [gf, gl], no two partial
    ranges share a slot, and the fix-up's slot rule finds each contributor's slot.  (The device code itself is
    covered by tests/test_gpu_kernels.py::test_streamk_schedule_matches_whole_tiles.)"""
    import random
    rnd = random.Random(1)
    max_contrib = 4
    for _ in range(3000):
        tiles, kb, slots = rnd.randint(1, 300), rnd.randint(16, 420), rnd.choice([74, 148, 66, 132])
        total = tiles * kb
        G = slots
        if total < G * 8:
            G = total // 8                                   # at least 8 k-blocks per CTA
        per_min = (kb - 1 + max_contrib - 2) // (max_contrib - 1)
        if G >= 1 and -(-total // G) < per_min:
            G = total // per_min                             # at most four CTAs per tile
        if G < 2:
            continue
        per = -(-total // G)
        G = -(-total // per)
        cover = {t: [] for t in range(tiles)}
        slot_of = {}
        for g in range(G):
            a, b = g * per, min(total, g * per + per)
            first_tile = (g * per) // kb
            while a < b:                                     # WorkIter<true>::next
                t = a // kb
                kb0 = a - t * kb
                kb1 = min(kb, kb0 + (b - a))
                a += kb1 - kb0
                partial = not (kb0 == 0 and kb1 == kb)
                if partial:
                    slot = g * 2 + (0 if t == first_tile else 1)
                    assert slot not in slot_of
                    slot_of[slot] = (t, g)
                cover[t].append((kb0, kb1, g, partial))
        for t in range(tiles):
            segs = sorted(cover[t])
            assert segs[0][0] == 0 and segs[-1][1] == kb
            assert all(x[1] == y[0] for x, y in zip(segs, segs[1:]))
            gf, gl = (t * kb) // per, ((t + 1) * kb - 1) // per   # igemm_fixup_kernel
            if gf == gl:
                assert len(segs) == 1 and not segs[0][3]
                continue
            assert [s[2] for s in segs] == list(range(gf, gl + 1)) and all(s[3] for s in segs)
            assert gl - gf + 1 <= max_contrib
            w_first = 0 if (gf * per) // kb == t else 1
            for _, _, g, _ in segs:
                assert slot_of[g * 2 + (w_first if g == gf else 0)] == (t, g)
