"""Kernel-level timeline of graph-replayed denoise steps (torch.profiler / CUPTI): busy time per
stream, idle gaps between consecutive kernels, overlap between the two streams."""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ctrlv_b200 import models, pipeline
T, h, w = 14, 40, 64
mu = models.UNetSpatioTemporalConditionModel(seed=0); mc = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=True)
st.set_schedule(sch.sigmas, sch.timesteps)
g = torch.Generator("cpu").manual_seed(1234)
st.latents.copy_(torch.randn(st.latents.shape, generator=g) * sch.init_noise_sigma)
st.capture()
for i in range(5): st.step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3): st.step(5 + i)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
print("cuda kernel events:", len(evs))
if not evs:
    sys.exit(0)
t0 = evs[0].time_range.start; t1 = max(e.time_range.end for e in evs)
span = (t1 - t0) / 3.0
busy_union = 0.0; cur_s, cur_e = None, None
for e in evs:
    s_, e_ = e.time_range.start, e.time_range.end
    if cur_e is None or s_ > cur_e:
        if cur_e is not None: busy_union += cur_e - cur_s
        cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy_union += cur_e - cur_s
tot = sum(e.time_range.end - e.time_range.start for e in evs)
byname = collections.defaultdict(lambda: [0, 0.0])
for e in evs:
    n = e.name.split("(")[0][-60:]
    byname[n][0] += 1; byname[n][1] += e.time_range.end - e.time_range.start
res = dict(kernels_per_step=len(evs) / 3, span_us_per_step=span, sum_kernel_us_per_step=tot / 3,
           busy_union_us_per_step=busy_union / 3, idle_us_per_step=span - busy_union / 3,
           overlap_us_per_step=(tot - busy_union) / 3)
print(json.dumps(res))
for n, (c, t) in sorted(byname.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{t/3/1e3:8.3f} ms n={c/3:6.0f} avg={t/c:7.1f} us  {n}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gap_profile.json", "w"), indent=1)
