"""Full-size (SVD config) parity on one B200: per-step latent rel-L2 of the sm_100a path vs the fp32
oracle over a 25-step trajectory, plus the torch-eager bf16 noise floor of the oracle modules."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
T0 = time.time()
def log(*a):
    print(f"[{time.time()-T0:7.1f}s]", *a, flush=True)
from oracle import svd_oracle as O
from oracle import sampling as S
from ctrlv_b200 import models, pipeline
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def main(steps=25, T=14, h=40, w=64):
    log("building weights on GPU (product random init), loading into oracle modules")
    cfg = dict(models.SVD_CONFIG)
    sd_u = models.random_state_dict(cfg, False, seed=0, dtype=torch.float32)
    sd_c = models.random_state_dict(cfg, True, seed=1, dtype=torch.float32, zero_conv_std=0.02)
    with torch.device("meta"):
        ou = O.UNetSpatioTemporalConditionModel(); oc = O.ControlNetModel()
    ou.load_state_dict(sd_u, assign=True); oc.load_state_dict(sd_c, assign=True)
    ou.eval(); oc.eval()
    log("oracle ready; packing product models")
    mu = models.UNetSpatioTemporalConditionModel(state_dict=sd_u)
    mc = models.ControlNetModel(state_dict=sd_c)
    log("packed")
    inp = S.make_inputs(T=T, h=h, w=w, device=dev)
    res = {}
    # --- single-step model-output error at a few sigmas
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
    for si in (0, 12, 22):
        sigma = sch.sigmas[si]; t = sch.timesteps[si]
        lat = inp["latents"] * float((sigma ** 2 + 1) ** 0.5)
        x = torch.cat([lat] * 2) / float((sigma ** 2 + 1) ** 0.5)
        x = torch.cat([x, inp["image_latents"]], dim=2)
        with torch.no_grad():
            od, om = oc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
            oy = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], od, om, return_dict=False)[0]
        md, mm = mc(x, t.to(dev), inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"], return_dict=False)
        my = mu(x, t.to(dev), inp["image_embeddings"], inp["added_time_ids"], md, mm, return_dict=False)[0]
        torch.cuda.synchronize()
        r = dict(ctrl_res_max=max(rel(a, b) for a, b in zip(md, od)), ctrl_mid=rel(mm, om), noise_pred=rel(my, oy))
        log(f"sigma={float(sigma):.4f}: {r}")
        res[f"single_step_sigma_{si}"] = r
        if si == 0:
            ub = ou.to(torch.bfloat16); cb = oc.to(torch.bfloat16)
            with torch.no_grad():
                bd, bm = cb(x.bfloat16(), t, inp["image_embeddings"].bfloat16(), inp["added_time_ids"].bfloat16(),
                            control_cond=inp["cond_em"].bfloat16(), return_dict=False)
                by = ub(x.bfloat16(), t, inp["image_embeddings"].bfloat16(), inp["added_time_ids"].bfloat16(), bd, bm, return_dict=False)[0]
            res["torch_bf16_noise_floor_noise_pred"] = rel(by, oy)
            log("torch-eager bf16 noise floor (noise_pred):", res["torch_bf16_noise_floor_noise_pred"])
            ou.load_state_dict(sd_u, assign=True); oc.load_state_dict(sd_c, assign=True)
            del ub, cb, bd, bm, by
    # --- trajectories
    trace = []
    t1 = time.time()
    with torch.no_grad():
        ofinal = S.sample_loop(ou, oc, inp, num_steps=steps, trace=trace)
    torch.cuda.synchronize()
    log(f"oracle fp32 GPU loop: {steps} steps in {time.time()-t1:.1f}s")
    pipe = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)
    mtrace = []
    out = pipe(cond_images=inp["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T, num_inference_steps=steps,
               latents=inp["latents"].clone(), output_type="latent", image_embeddings=inp["image_embeds_cond"],
               image_latents=inp["image_latents_cond"],
               callback_on_step_end=lambda p, i, t, kw: mtrace.append(kw["latents"].clone()) or {})
    torch.cuda.synchronize()
    rs = [rel(a, b) for a, b in zip(mtrace, trace)]
    res["trajectory_per_step_rel_l2"] = rs
    res["final_rel_l2"] = rel(out.frames, ofinal)
    log("trajectory per-step latent rel_l2:", [f"{r:.2e}" for r in rs])
    log("final:", res["final_rel_l2"], "max:", max(rs))
    # --- one-step-from-oracle-state (teacher forced) per-step error
    st = next(iter(pipe._steps.values()))
    tf = []
    prev = inp["latents"] * float(sch.init_noise_sigma)
    sch2 = S.EulerDiscreteSchedulerOracle(); sch2.set_timesteps(steps)
    prevs = [inp["latents"].to(dev) * sch2.init_noise_sigma] + trace[:-1]
    for i in range(steps):
        st.latents.copy_(prevs[i])
        st.step(i)
        tf.append(rel(st.latents, trace[i]))
    res["teacher_forced_per_step_rel_l2"] = tf
    log("teacher-forced per-step latent rel_l2:", [f"{r:.2e}" for r in tf], "max", max(tf))
    # --- decoded-frame PSNR after the 25 steps (north-star criterion): full-size temporal VAE, random
    # init; new chain (sm_100a loop + sm_100a decode) vs oracle chain (fp32 loop + fp32 decode)
    import math
    from oracle import vae_oracle as V
    from ctrlv_b200 import vae
    del ou, oc
    torch.cuda.empty_cache()
    torch.manual_seed(3)
    ov = V.AutoencoderKLTemporalDecoder().to(dev).eval()
    mv = vae.AutoencoderKLTemporalDecoder(state_dict=ov.state_dict())
    # random-init UNets do not denoise: bring the final latents to the scale a VAE expects so that the
    # decode is exercised in its working range (same factor on both chains)
    k = float(0.18215 / ofinal.std())
    with torch.no_grad():
        want = V.decode_latents(ov, ofinal * k, T, T)
    got = vae.decode_latents(mv, out.frames * k, T, T)
    to01 = lambda v: (v / 2 + 0.5).clamp(0, 1)
    mse = float(((to01(got) - to01(want)) ** 2).mean())
    mse_raw = float(((got - want) ** 2).mean()); peak = float(want.max() - want.min())
    res["decoded_psnr_db_clamped_0_1"] = 10 * math.log10(1.0 / max(mse, 1e-30))
    res["decoded_psnr_db_oracle_range"] = 10 * math.log10(peak * peak / max(mse_raw, 1e-30))
    res["decoded_rel_l2"] = rel(got, want)
    log("decoded frames", list(got.shape), "PSNR [0,1]-clamped:", res["decoded_psnr_db_clamped_0_1"],
        "PSNR oracle-range:", res["decoded_psnr_db_oracle_range"], "rel_l2:", res["decoded_rel_l2"])
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/parity_full.json", "w"), indent=1)
    # quick timing of the graph step
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for i in range(3): st.step(i)
    e0.record()
    for i in range(10): st.step(i)
    e1.record(); torch.cuda.synchronize()
    log(f"graph step: {e0.elapsed_time(e1)/10:.2f} ms/step")

if __name__ == "__main__":
    main()
