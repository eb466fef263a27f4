"""'Stock GPU' line: the oracle modules (restated reference math) in torch-eager bf16 on one B200
(cuDNN / cuBLAS / SDPA), same CFG denoise step, timed with CUDA events."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import svd_oracle as O
from oracle import sampling as S
from ctrlv_b200 import models
dev = "cuda"
T = int(os.environ.get("T", 14)); h = int(os.environ.get("H", 40)); w = int(os.environ.get("W", 64))
cfg = dict(models.SVD_CONFIG)
sd_u = models.random_state_dict(cfg, False, seed=0, dtype=torch.bfloat16)
sd_c = models.random_state_dict(cfg, True, seed=1, dtype=torch.bfloat16, zero_conv_std=0.02)
with torch.device("meta"):
    ou = O.UNetSpatioTemporalConditionModel(); oc = O.ControlNetModel()
ou.load_state_dict(sd_u, assign=True); oc.load_state_dict(sd_c, assign=True)
inp = {k: (v.to(torch.bfloat16) if v.is_floating_point() else v) for k, v in S.make_inputs(T=T, h=h, w=w, device=dev).items()}
sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
gs = inp["guidance"].view(1, -1, 1, 1, 1)
lat = inp["latents"] * sch.init_noise_sigma
def step(i):
    sch.step_index = i
    with torch.no_grad():
        return S.denoise_step(ou, oc, sch, lat, sch.timesteps[i], inp["image_latents"], inp["image_embeddings"],
                              inp["added_time_ids"], inp["cond_em"], gs)
for i in range(3): step(i)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for i in range(n): step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"impl": "torch-eager bf16 oracle modules (cuDNN/cuBLAS/SDPA)", "T": T, "latent": [h, w],
                  "ms_per_step": ms, "steps_per_s": 1e3 / ms}))
