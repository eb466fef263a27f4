"""Generates tests/golden/tiny_step.pt from the oracle (CPU, fp32).

The reference ships no golden vectors and its arithmetic lives in the absent
diffusers==0.27.2, so these fixtures pin the ORACLE RESTATEMENT against drift, not the reference:
that part of the parity stays "unpinned" in the sense of the task statement (the in-repo reference code is
pinned separately: make_ref_golden.py / ref_forward.pt).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import svd_oracle as O  # noqa: E402
from oracle import sampling as S  # noqa: E402

T, H, W = 2, 8, 8


def build():
    torch.manual_seed(0)
    unet = O.UNetSpatioTemporalConditionModel(**O.TINY_CONFIG).eval()
    ctrl = O.ControlNetModel(**O.TINY_CONFIG).eval()
    O.randomize_zero_convs(ctrl)
    return unet, ctrl


def run():
    unet, ctrl = build()
    inp = S.make_inputs(T=T, h=H, w=W, xdim=O.TINY_CONFIG["cross_attention_dim"])
    sch = S.EulerDiscreteSchedulerOracle()
    sch.set_timesteps(25)
    out = {"T": T, "H": H, "W": W, "sigmas": sch.sigmas.clone(), "timesteps": sch.timesteps.clone(),
           "init_noise_sigma": float(sch.init_noise_sigma)}
    with torch.no_grad():
        lat = inp["latents"] * sch.init_noise_sigma
        gs = inp["guidance"].view(1, -1, 1, 1, 1)
        lat1, noise = S.denoise_step(unet, ctrl, sch, lat, sch.timesteps[0], inp["image_latents"],
                                     inp["image_embeddings"], inp["added_time_ids"], inp["cond_em"], gs,
                                     return_noise=True)
        x = torch.cat([lat] * 2) / ((sch.sigmas[0] ** 2 + 1) ** 0.5)
        x = torch.cat([x, inp["image_latents"]], dim=2)
        down, mid = ctrl(x, sch.timesteps[0], inp["image_embeddings"], inp["added_time_ids"],
                         control_cond=inp["cond_em"], return_dict=False)
    out["noise_pred"] = noise
    out["latents_after_step0"] = lat1
    out["ctrl_down_norms"] = torch.stack([d.norm() for d in down])
    out["ctrl_mid"] = mid
    out["unet_param_checksum"] = float(sum(p.detach().double().sum() for p in unet.parameters()))
    return out


if __name__ == "__main__":
    o = run()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_step.pt")
    torch.save(o, path)
    print("wrote", path, os.path.getsize(path), "bytes")
