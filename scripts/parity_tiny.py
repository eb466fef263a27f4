"""GPU parity of the drop-in models / fused step against the fp32 oracle (reduced config)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import svd_oracle as O
from oracle import sampling as S
import ctrlv_b200
from ctrlv_b200 import models, pipeline

dev = "cuda"

def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def main(cfg_name="tiny", T=4, h=16, w=16, order="s_major", steps=3):
    over = dict(O.TINY_CONFIG) if cfg_name == "tiny" else {}
    xdim = over.get("cross_attention_dim", 1024)
    torch.manual_seed(0)
    t0 = time.time()
    ou = O.UNetSpatioTemporalConditionModel(time_context_order=order, **over)
    oc = O.ControlNetModel(time_context_order=order, **over)
    O.randomize_zero_convs(oc)
    ou = ou.to(dev).eval(); oc = oc.to(dev).eval()
    print(f"[{cfg_name} {order}] oracle built in {time.time()-t0:.1f}s", flush=True)
    mu = models.UNetSpatioTemporalConditionModel(state_dict=ou.state_dict(), time_context_order=order, **over)
    mc = models.ControlNetModel(state_dict=oc.state_dict(), time_context_order=order, **over)
    inp = S.make_inputs(T=T, h=h, w=w, xdim=xdim, device=dev)
    sch = S.EulerDiscreteSchedulerOracle(); sch.set_timesteps(25)
    ok = True
    for si in (0, 12, 24):
        sigma = sch.sigmas[si]; t = sch.timesteps[si]
        lat = inp["latents"] * float(sch.init_noise_sigma) if si == 0 else inp["latents"] * float(sigma)
        x = torch.cat([lat] * 2) / float((sigma ** 2 + 1) ** 0.5)
        x = torch.cat([x, inp["image_latents"]], dim=2)
        with torch.no_grad():
            od, om = oc(x, t, inp["image_embeddings"], inp["added_time_ids"], control_cond=inp["cond_em"],
                        conditioning_scale=1.0, return_dict=False)
            oy = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], od, om, return_dict=False)[0]
            oy0 = ou(x, t, inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
        md, mm = mc(x, timestep=t.to(dev), encoder_hidden_states=inp["image_embeddings"],
                    added_time_ids=inp["added_time_ids"], control_cond=inp["cond_em"], conditioning_scale=1.0,
                    return_dict=False)
        my = mu(x, t.to(dev), inp["image_embeddings"], inp["added_time_ids"], md, mm, return_dict=False)[0]
        my0 = mu(x, t.to(dev), inp["image_embeddings"], inp["added_time_ids"], return_dict=False)[0]
        torch.cuda.synchronize()
        rd = [rel(a, b) for a, b in zip(md, od)]
        print(f" step {si} sigma={float(sigma):.4f}: ctrl residual rel_l2 max={max(rd):.3e} mid={rel(mm, om):.3e} | "
              f"unet+ctrl={rel(my, oy):.3e} unet only={rel(my0, oy0):.3e} | ctrl effect={rel(oy, oy0):.3e}", flush=True)
        ok &= max(rd) < 2e-2 and rel(my, oy) < 2e-2
        # torch-eager bf16 noise floor of the same modules
        if si == 0:
            bu = O.UNetSpatioTemporalConditionModel(time_context_order=order, **over); bu.load_state_dict(ou.state_dict()); bu = bu.to(dev, torch.bfloat16)
            bc = O.ControlNetModel(time_context_order=order, **over); bc.load_state_dict(oc.state_dict()); bc = bc.to(dev, torch.bfloat16)
            with torch.no_grad():
                xb = x.to(torch.bfloat16)
                bd, bm = bc(xb, t, inp["image_embeddings"].bfloat16(), inp["added_time_ids"].bfloat16(),
                            control_cond=inp["cond_em"].bfloat16(), return_dict=False)
                by = bu(xb, t, inp["image_embeddings"].bfloat16(), inp["added_time_ids"].bfloat16(), bd, bm, return_dict=False)[0]
            print(f"   torch-eager bf16 noise floor vs fp32 oracle: unet+ctrl={rel(by, oy):.3e}", flush=True)
            del bu, bc
    # fused step + loop vs oracle loop
    trace = []
    with torch.no_grad():
        ofinal = S.sample_loop(ou, oc, inp, num_steps=steps, trace=trace)
    pipe = pipeline.StableVideoControlPipeline(unet=mu, controlnet=mc)
    mtrace = []
    for use_graph in (False, True):
        mtrace.clear()
        out = pipe(cond_images=inp["cond_em_cond"], height=h * 8, width=w * 8, num_frames=T,
                   num_inference_steps=steps, latents=inp["latents"].clone(), output_type="latent",
                   image_embeddings=inp["image_embeds_cond"], image_latents=inp["image_latents_cond"],
                   use_graph=use_graph,
                   callback_on_step_end=lambda p, i, t, kw: mtrace.append(kw["latents"].clone()) or {})
        torch.cuda.synchronize()
        rs = [rel(a, b) for a, b in zip(mtrace, trace)]
        print(f" pipeline loop graph={use_graph}: per-step latent rel_l2 = {[f'{r:.2e}' for r in rs]} final={rel(out.frames, ofinal):.3e}", flush=True)
        ok &= max(rs) < 1e-2
    print("PARITY OK" if ok else "PARITY FAIL", flush=True)
    return ok

if __name__ == "__main__":
    ok = main("tiny", 4, 16, 16, "s_major")
    ok &= main("tiny", 5, 24, 40, "b_major", steps=2)
    if "--full" in sys.argv:
        ok &= main("full", 14, 40, 64, "s_major", steps=2)
    sys.exit(0 if ok else 1)
