"""One eager (no CUDA graph) denoise step between cudaProfilerStart/Stop, for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, pipeline

T = int(os.environ.get("T", 14)); h = int(os.environ.get("H", 40)); w = int(os.environ.get("W", 64))
mu = models.UNetSpatioTemporalConditionModel(seed=0)
mc = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=False)
st.set_schedule(sch.sigmas, sch.timesteps)
g = torch.Generator("cpu").manual_seed(1234)
st.latents.copy_(torch.randn(st.latents.shape, generator=g) * sch.init_noise_sigma)
st.image_latents[1].copy_(torch.randn(st.image_latents[1].shape, generator=g))
st.cond_em[1].copy_(torch.randn(st.cond_em[1].shape, generator=g))
st.ehs[1].copy_(torch.randn(st.ehs[1].shape, generator=g))
st.added_time_ids.copy_(torch.tensor([[6.0, 127.0, 0.02]] * 2))
st.guidance.copy_(torch.linspace(1, 3, T))
st.capture()   # eager warm-up only (use_graph=False)
st.step(1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.step(2)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
