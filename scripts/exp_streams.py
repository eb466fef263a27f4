import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, pipeline
mu = models.UNetSpatioTemporalConditionModel(seed=0); mc = models.ControlNetModel(seed=1)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
res = {}
for two in (False, True, False, True):
    st = pipeline.DenoiseStep(mu, mc, 1, 14, 40, 64, cfg=True, use_graph=True, two_streams=two)
    st.set_schedule(sch.sigmas, sch.timesteps)
    g = torch.Generator("cpu").manual_seed(1234)
    st.latents.copy_(torch.randn(st.latents.shape, generator=g) * sch.init_noise_sigma)
    st.capture()
    for i in range(3): st.step(i)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(25): st.step(i)
    e1.record(); torch.cuda.synchronize()
    print(f"two_streams={two}: {e0.elapsed_time(e1)/25:.2f} ms/step", flush=True)
    res[two] = st.latents.clone()
    del st
print("bitwise equal:", torch.equal(res[False], res[True]))
