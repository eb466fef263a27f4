"""BASELINE config 3 on ONE GPU: 8 clips x 25 Euler-EDM steps with CFG, as whole clips per GPU
(SURVEY.md §8e).  Compares running the clips one at a time (CFG batch 2, the bench.py step) with
packing 2 or 4 clips into one step (CFG batch 4 / 8): same kernels, larger M per launch.
Writes gpurun_out/config3.json.  Not a bench.py line: the headline metric stays batch 1."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, pipeline

T, h, w, CLIPS, STEPS = 14, 40, 64, 8, 25
batches = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2", "4"])]
mu = models.UNetSpatioTemporalConditionModel(seed=0)
mc = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(STEPS)
res = {"config": "8 clips x 25 steps, 14x320x512, CFG, one B200", "runs": []}
for b in batches:
    st = pipeline.DenoiseStep(mu, mc, b, T, h, w, cfg=True, use_graph=True)
    st.set_schedule(sch.sigmas, sch.timesteps)
    # synthetic inputs of SURVEY §8d (seed 1234, drawn on the CPU in fp32, uncond halves zero)
    g = torch.Generator("cpu").manual_seed(1234)
    lat0 = (torch.randn(b, T, 4, h, w, generator=g) * sch.init_noise_sigma).cuda()
    il = torch.randn(b, 4, h, w, generator=g).unsqueeze(1).repeat(1, T, 1, 1, 1)
    emb = torch.randn(b, 1024, generator=g)
    cond = torch.randn(b, T, 4, h, w, generator=g)
    st.image_latents.copy_(torch.cat([torch.zeros_like(il), il])); st.cond_em.copy_(torch.cat([torch.zeros_like(cond), cond]))
    st.ehs.copy_(torch.cat([torch.zeros_like(emb), emb])); st.added_time_ids.copy_(torch.tensor([[6.0, 127.0, 0.02]] * (2 * b)))
    st.guidance.copy_(torch.linspace(1.0, 3.0, T))
    st.latents.copy_(lat0); st.capture()
    for k in range(3):
        st.step(k)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(CLIPS // b):          # the whole job: every group of b clips, all 25 steps
        st.latents.copy_(lat0)
        for k in range(STEPS):
            st.step(k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    r = {"clips_per_step": b, "cfg_batch": 2 * b, "job_ms": ms, "ms_per_step": ms / (STEPS * CLIPS // b),
         "ms_per_clip_step": ms / (STEPS * CLIPS), "clips_per_min": CLIPS * 60e3 / ms,
         "finite": bool(torch.isfinite(st.latents).all())}
    res["runs"].append(r)
    print(json.dumps(r), flush=True)
    del st
    torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/config3.json", "w"), indent=1)
