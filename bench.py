"""bench.py — Box2Video denoise-step throughput on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N --steps K --warmup W   # reference CPU path (oracle)

A "step" is one Euler-EDM iteration of ControlNet + UNet under classifier-free guidance
(pipeline_video_control.py:298-343) on one clip of 14 frames at 320x512 (latent 14x4x40x64, CFG
batch 2), bf16, synthetic inputs, random-init weights of the full SVD architecture
(BASELINE.json configs[1]).  For N > 1 every rank samples its own clip (sample-parallel, weak
scaling); value = steps of all ranks / max-over-ranks device time.  One JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "denoise_steps_per_s"
UNIT = "steps/s"
STEP_TFLOP = 29.14  # algorithmic work of one CFG step at 14x320x512


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--height", type=int, default=320)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU-baseline sample (0 = auto)")
    return ap.parse_args()


def workload_name(a):
    return (f"Box2Video CFG denoise step: SVD UNet (1.52B) + bbox ControlNet (0.68B), 1 clip x {a.frames} frames "
            f"@{a.height}x{a.width} (latent {a.frames}x4x{a.height // 8}x{a.width // 8}), 25-step Euler-EDM schedule")


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference path (the oracle restatement of the reference math, fp32, all host cores)
# ------------------------------------------------------------------------------------------------
def build_oracle_cpu():
    import torch
    from oracle import svd_oracle as O
    with torch.device("meta"):
        ou = O.UNetSpatioTemporalConditionModel()
        oc = O.ControlNetModel()
    g = torch.Generator().manual_seed(0)

    def fill(mod):
        sd = {}
        for k, v in mod.state_dict().items():
            if v.dim() == 1 and (".norm" in k or "conv_norm_out" in k) and k.endswith("weight"):
                sd[k] = torch.ones(v.shape)
            elif v.dim() == 1 and (".norm" in k or "conv_norm_out" in k):
                sd[k] = torch.zeros(v.shape)
            elif k.endswith("mix_factor"):
                sd[k] = torch.full(v.shape, 0.5)
            else:
                fan = max(1, v[0].numel()) if v.dim() > 1 else 64
                sd[k] = (torch.rand(v.shape, generator=g) * 2 - 1) / fan ** 0.5
        mod.load_state_dict(sd, assign=True)
        return mod.eval()
    return fill(ou), fill(oc)


def cpu_step_time(ou, oc, frames, h, w, repeats=1):
    """Seconds for one CFG denoise step of the oracle on the host cores with `frames` frames."""
    import torch
    from oracle import sampling as S
    inp = S.make_inputs(T=frames, h=h, w=w)
    sch = S.EulerDiscreteSchedulerOracle()
    sch.set_timesteps(25)
    lat = inp["latents"] * sch.init_noise_sigma
    gs = inp["guidance"].view(1, -1, 1, 1, 1)
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            sch.step_index = 0
            t0 = time.time()
            S.denoise_step(ou, oc, sch, lat, sch.timesteps[0], inp["image_latents"], inp["image_embeddings"],
                           inp["added_time_ids"], inp["cond_em"], gs)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
    return best


def cpu_baseline(a, budget_s=25.0):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ou, oc = build_oracle_cpu()
    h, w = a.height // 8, a.width // 8
    frames = a.cpu_frames
    if frames <= 0:
        t1 = cpu_step_time(ou, oc, 1, h, w)  # probe (also warms the thread pool / allocator)
        frames = max(1, min(a.frames, int(budget_s / max(t1, 1e-3))))
    dt = cpu_step_time(ou, oc, frames, h, w)
    scale = frames / a.frames
    return {"value": scale / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": (f"oracle fp32 (torch CPU, {cores} threads): 1 CFG step of ControlNet+UNet on {frames} of "
                       f"{a.frames} frames @{a.height}x{a.width} in {dt:.1f} s, scaled by {frames}/{a.frames} to the full step")}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ou, oc = build_oracle_cpu()
    h, w = a.height // 8, a.width // 8
    n = a.steps + a.warmup
    t1 = cpu_step_time(ou, oc, 1, h, w)
    budget = 150.0 / max(n, 1)  # keep the whole run within a few minutes
    frames = max(1, min(a.frames, int(budget / max(t1, 1e-3))))
    for _ in range(a.warmup):
        cpu_step_time(ou, oc, frames, h, w)
    t0 = time.time()
    for _ in range(a.steps):
        cpu_step_time(ou, oc, frames, h, w)
    dt = (time.time() - t0) / a.steps
    value = (frames / a.frames) / dt
    sample = (f"oracle fp32 (torch CPU, {cores} threads): each step = 1 CFG step of ControlNet+UNet on {frames} of "
              f"{a.frames} frames @{a.height}x{a.width} ({dt:.2f} s), scaled by {frames}/{a.frames}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "reference_path": "oracle/ (diffusers==0.27.2 restatement; "
                       "the reference itself cannot be installed offline)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from ctrlv_b200 import _lib, models, ops, parallel, pipeline

    rank, world, local = parallel.init_from_env("nccl")
    torch.cuda.set_device(local)
    lib = _lib.load(build_if_missing=False)  # the product path: no fallback
    if lib.ctrlv_device_check() != 0:
        raise RuntimeError(lib.ctrlv_last_error().decode())

    T, h, w = a.frames, a.height // 8, a.width // 8
    unet = models.UNetSpatioTemporalConditionModel(seed=0)
    ctrl = models.ControlNetModel(seed=1)
    sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
    st = pipeline.DenoiseStep(unet, ctrl, 1, T, h, w, cfg=True, conditioning_scale=1.0, use_graph=True)
    st.set_schedule(sch.sigmas, sch.timesteps)
    # synthetic inputs: seed 1234 + rank, drawn on the CPU in fp32
    g = torch.Generator("cpu").manual_seed(1234 + rank)
    hb = st.make_host_buffers()
    hb["latents"].copy_(torch.randn(st.latents.shape, generator=g) * sch.init_noise_sigma)
    il = torch.randn(1, 4, h, w, generator=g)
    hb["image_latents"].zero_(); hb["image_latents"][1:].copy_(il.unsqueeze(1).expand(1, T, 4, h, w))
    hb["ehs"].zero_(); hb["ehs"][1:].copy_(torch.randn(1, st.ehs.shape[1], generator=g))
    hb["cond_em"].zero_(); hb["cond_em"][1:].copy_(torch.randn(1, T, 4, h, w, generator=g))
    hb["added_time_ids"].copy_(torch.tensor([[6.0, 127.0, 0.02]] * 2))
    hb["guidance"].copy_(torch.linspace(1.0, 3.0, T))
    for k in st.HOST_INPUTS:
        getattr(st, k).copy_(hb[k])
    st.capture()
    nsched = sch.timesteps.numel()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # count this repo's kernel launches per step (eager pass with the per-call hook)
    ops.PROFILE = {}
    st._graph, gsave = None, st._graph
    lat_keep = st.latents.clone()
    st.step(0)
    ops.profile_flush()
    prof = ops.PROFILE
    ops.PROFILE = None
    st._graph = gsave
    st.latents.copy_(lat_keep)
    # launches: every profiled op is one kernel except groupnorm (2); plus the un-profiled glue ops
    n_prof = sum(v[0] * {"groupnorm": 2, "upconv3x3": 4}.get(k[0], 1) for k, v in prof.items())
    glue = 2 + 2 * (2 + 4 + 2) + 13  # prep + cfg_euler, per-model embeds (sinusoid x2, MLP x4, aux x2), axpby
    launches_per_step = int(n_prof + glue)
    gemm_ms = sum(v[1] for k, v in prof.items() if k[0] in ("linear", "conv3x3", "conv_t3", "upconv3x3"))
    gemm_fl = sum(v[2] for k, v in prof.items() if k[0] in ("linear", "conv3x3", "conv_t3", "upconv3x3"))
    gemm_n = sum(v[0] * (4 if k[0] == "upconv3x3" else 1) for k, v in prof.items()
                 if k[0] in ("linear", "conv3x3", "conv_t3", "upconv3x3"))

    # ---- device-resident timing: W warm-up steps, then exactly K steps
    for i in range(a.warmup):
        st.step(i % nsched)
    if world > 1:  # warm-up of the job's one collective too (NCCL connects its all-gather channels lazily)
        parallel.gather_latents(st.latents, world, world, rank)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        st.step(i % nsched)
    if world > 1:  # collect the clips' latents (the path's only collective)
        parallel.gather_latents(st.latents, world, world, rank)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    barrier()
    clk = clocks.stop()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = world * a.steps / (ms / 1e3)

    # ---- end to end: host buffers in, host latents out, every step
    for i in range(min(a.warmup, 3)):
        st.step_host(i % nsched, hb)
    barrier()
    e0.record()
    for i in range(a.steps):
        st.step_host(i % nsched, hb)
        hb["latents"].copy_(hb["latents_out"])  # next step consumes the result, like the reference loop
    e1.record()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    barrier()
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = world * a.steps / (float(ms2.item()) / 1e3)
    h2d, d2h = st.host_bytes_per_step()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return 0
    peaks, peak_src = {}, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM bytes per igemm launch (average over the step's 429 launches) from the committed ncu pass
    traffic = None
    if (T, h, w) == (14, 40, 64):
        try:
            traffic = float(json.load(open(os.path.join(ROOT, "profiles", "r01_step_dram_traffic.json")))
                            ["igemm_dram_bytes_per_launch"])
        except Exception:
            traffic = None
    roofline = {"bound": "tensor", "kernel": "ctrlv::igemm_kernel (tcgen05 implicit GEMM: Linear / 3x3 / temporal conv)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "profiles/r01_step_dram_traffic.json: ncu dram__bytes_read+write, cold caches, mean per igemm launch",
                "peak_source": peak_src, "launches_per_step": int(gemm_n),
                "timing": "CUDA events around every igemm launch of one eager (un-captured) step in this process, on the "
                          "launching stream; the timed region replays the same launches from a CUDA graph, which "
                          "cannot be event-timed per kernel",
                "algorithmic_tflop_per_step": gemm_fl / 1e12, "avg_launch_us": 1e3 * gemm_ms / max(gemm_n, 1),
                "whole_step_tflops": STEP_TFLOP * value / world if (T, h, w) == (14, 40, 64) else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a), "clips_per_gpu": 1, "cfg_batch": 2, "parallelism": f"sample-parallel x{world}",
                       "weights": "random-init, full SVD architecture", "l2_policy": "inputs+weights (4.4 GB bf16) exceed the 126 MB L2; no flush",
                       "clips_per_min_at_25_steps": value * 60.0 / 25.0, "cuda_graph": True},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * a.steps,
            "roofline": roofline}
    if world == 1 and not a.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(a)
    print(json.dumps(line), flush=True)
    return 0


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    return run_b200(a)


if __name__ == "__main__":
    sys.exit(main())
