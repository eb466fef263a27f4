"""Per-op / per-shape timing table of one eager denoise step (CUDA events around each call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, pipeline, ops
T = int(os.environ.get("T", 14)); h = int(os.environ.get("H", 40)); w = int(os.environ.get("W", 64))
mu = models.UNetSpatioTemporalConditionModel(seed=0)
mc = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=False)
st.set_schedule(sch.sigmas, sch.timesteps)
g = torch.Generator("cpu").manual_seed(1234)
st.latents.copy_(torch.randn(st.latents.shape, generator=g) * sch.init_noise_sigma)
st.capture(); st.step(1); torch.cuda.synchronize()
ops.PROFILE = {}
for i in range(3): st.step(2 + i)
ops.profile_flush()
tot = sum(v[1] for v in ops.PROFILE.values()) / 3
print(f"sum of timed ops: {tot:.2f} ms/step")
byop = {}
for (op, key), (n, ms, fl, _by) in ops.PROFILE.items():
    r = byop.setdefault(op, [0, 0.0, 0.0]); r[0] += n / 3; r[1] += ms / 3; r[2] += fl / 3
for op, (n, ms, fl) in sorted(byop.items(), key=lambda kv: -kv[1][1]):
    print(f"{op:14s} n={n:5.0f} {ms:8.3f} ms  {fl/ms/1e9 if fl else 0:8.1f} TFLOP/s")
print("--- per shape (sorted by time)")
for (op, key), (n, ms, fl, _by) in sorted(ops.PROFILE.items(), key=lambda kv: -kv[1][1]):
    n /= 3; ms /= 3; fl /= 3
    print(f"{ms:8.3f} ms n={n:4.0f} avg={1e3*ms/n:8.1f} us {fl/ms/1e9 if fl else 0:7.1f} TF  {op} {key}")
