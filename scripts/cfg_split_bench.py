"""CFG-branch-sharded sampling of ONE clip on a GPU pair: uncond on rank 2k, cond on
rank 2k+1, one 573 KB all-gather of the model output per step.  Reports the per-step latency next to
the batched (one GPU, both branches) step and checks the two against each other.
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/cfg_split_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ctrlv_b200 import models, parallel, pipeline

rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
T, h, w, steps = 14, 40, 64, 25
unet = models.UNetSpatioTemporalConditionModel(seed=0)
ctrl = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
pair = parallel.CfgPair(rank, world)
g = torch.Generator("cpu").manual_seed(1234 + pair.pair)


def fill(st):
    gg = torch.Generator("cpu").manual_seed(1234 + pair.pair)
    st.set_schedule(sch.sigmas, sch.timesteps)
    lat = torch.randn(st.latents.shape, generator=gg) * sch.init_noise_sigma
    st.image_latents.zero_(); st.image_latents[1:].copy_(torch.randn(1, 4, h, w, generator=gg).unsqueeze(1).expand(1, T, 4, h, w))
    st.ehs.zero_(); st.ehs[1:].copy_(torch.randn(1, st.ehs.shape[1], generator=gg))
    st.cond_em.zero_(); st.cond_em[1:].copy_(torch.randn(1, T, 4, h, w, generator=gg))
    st.added_time_ids.copy_(torch.tensor([[6.0, 127.0, 0.02]] * 2)); st.guidance.copy_(torch.linspace(1.0, 3.0, T))
    st.capture()
    st.latents.copy_(lat)
    return lat


def timed(st, n):
    for i in range(3):
        st.step(i)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        st.step(i % 25)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


split = pipeline.DenoiseStep(unet, ctrl, 1, T, h, w, cfg=True, cfg_branch=pair.branch, exchange=pair.exchange)
lat0 = fill(split)
ms_split = timed(split, steps)
full = pipeline.DenoiseStep(unet, ctrl, 1, T, h, w, cfg=True)
fill(full)
ms_full = timed(full, steps)
# parity: a whole 25-step trajectory, split vs batched, from the same start
split.latents.copy_(lat0); full.latents.copy_(lat0)
for i in range(25):
    split.step(i); full.step(i)
torch.cuda.synchronize()
rel = float((split.latents - full.latents).norm() / full.latents.norm())
same = bool(torch.equal(split.latents, full.latents))
if rank == 0:
    print(json.dumps({"n_gpus": world, "pairs": world // 2, "ms_per_step_cfg_split_pair": ms_split,
                      "ms_per_step_batched_one_gpu": ms_full, "latency_speedup": ms_full / ms_split,
                      "trajectory_rel_l2_split_vs_batched": rel, "bitwise_equal": same}), flush=True)
dist.barrier()
dist.destroy_process_group()
