"""BASELINE config 4: SVD-XT shape — 25 frames at 576x1024 (latent 25x4x72x128), one CFG step:
timing of the sm_100a path + teacher-forced parity against the fp32 oracle on the GPU."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import svd_oracle as O
from oracle import sampling as S
from ctrlv_b200 import models, pipeline
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
dev = "cuda"; T, h, w = 25, 72, 128
cfg = dict(models.SVD_CONFIG)
sd_u = models.random_state_dict(cfg, False, seed=0, dtype=torch.float32)
sd_c = models.random_state_dict(cfg, True, seed=1, dtype=torch.float32, zero_conv_std=0.02)
mu = models.UNetSpatioTemporalConditionModel(state_dict=sd_u); mc = models.ControlNetModel(state_dict=sd_c)
inp = S.make_inputs(T=T, h=h, w=w, device=dev)
sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=True)
st.set_schedule(sch.sigmas, sch.timesteps)
st.image_latents.copy_(inp["image_latents"]); st.cond_em.copy_(inp["cond_em"]); st.ehs.copy_(inp["image_embeddings"].reshape(2, -1))
st.added_time_ids.copy_(inp["added_time_ids"]); st.guidance.copy_(inp["guidance"])
i = 12
lat = inp["latents"] * float((sch.sigmas[i] ** 2 + 1) ** 0.5)
st.latents.copy_(lat); st.capture()
st.latents.copy_(lat); st.step(i); torch.cuda.synchronize()
got = st.latents.clone(); got_noise = st.noise.clone()
for k in range(3): st.step(k)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(10): st.step(k)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
res = {"config": "SVD-XT 25x576x1024 (latent 25x4x72x128), CFG batch 2", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
       "tflops_reference_math": 218.5 / ms}
print(json.dumps(res), flush=True)
del st
with torch.device("meta"):
    ou = O.UNetSpatioTemporalConditionModel(); oc = O.ControlNetModel()
ou.load_state_dict(sd_u, assign=True); oc.load_state_dict(sd_c, assign=True)
sch.step_index = i
t0 = time.time()
with torch.no_grad():
    want, noise = S.denoise_step(ou, oc, sch, lat, sch.timesteps[i], inp["image_latents"], inp["image_embeddings"],
                                 inp["added_time_ids"], inp["cond_em"], inp["guidance"].view(1, -1, 1, 1, 1), return_noise=True)
torch.cuda.synchronize()
rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm())
res["oracle_fp32_gpu_s"] = time.time() - t0
res["latent_rel_l2"] = rel(got, want)
res["noise_pred_rel_l2"] = rel(got_noise.view(2, T, h, w, 4).permute(0, 1, 4, 2, 3), noise)
print(json.dumps(res), flush=True)
json.dump(res, open("gpurun_out/config4.json", "w"), indent=1)
