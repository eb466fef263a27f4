"""Record every distinct igemm problem of one denoise step (ops.PROFILE keys), then sweep n-tile width
and cta_group for each with synthetic operands: auto choice vs best, weighted by launches per step."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import models, pipeline, ops
BF = torch.bfloat16; dev = "cuda"
T, h, w = 14, 40, 64
mu = models.UNetSpatioTemporalConditionModel(seed=0); mc = models.ControlNetModel(seed=1, zero_conv_std=0.02)
sch = pipeline.EulerDiscreteScheduler().set_timesteps(25)
st = pipeline.DenoiseStep(mu, mc, 1, T, h, w, cfg=True, use_graph=False, two_streams=False)
st.set_schedule(sch.sigmas, sch.timesteps)
st.capture()
ops.PROFILE = {}
st.step(1); ops.profile_flush()
keys = {k: v[0] for k, v in ops.PROFILE.items() if k[0] in ("linear", "conv3x3", "conv_t3", "upconv3x3")}
ops.PROFILE = None
del st, mu, mc
torch.cuda.empty_cache()
torch.manual_seed(0)
big = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, n=5):
    fn(); fn()
    ts = []
    for _ in range(n):
        big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]
def rnd(*s): return torch.randn(*s, device=dev).to(BF)
def gn_arg(flag, M, N):
    if not flag or not ops.GNStats.fusable(N):
        return None
    units = 28 if M % 28 == 0 else 1
    return (ops.GNStats(torch.zeros(ops.GNStats.numel(units), dtype=torch.int64, device=dev), units, M // units, N), 0)
def make(op, key):
    if op == "linear":
        M, K, N, geglu, r1, r2, gn = key
        a = rnd(M, K); wt = (torch.randn(N, K, device=dev) / K ** 0.5).to(BF); b = torch.randn(N, device=dev)
        No = N // 2 if geglu else N
        kw = dict(bias=b, out=torch.empty(M, No, device=dev, dtype=BF), geglu=geglu)
        if r1: kw["res1"] = rnd(M, No)
        if r2: kw["res2"] = rnd(M, No)
        kw["gn"] = gn_arg(gn, M, No)
        return (lambda: ops.linear(a, wt, **kw)), N
    if op == "conv3x3":
        F_, H, W, stride, Cin, SC, N, gn = key
        x = rnd(F_ * H * W, Cin); kk = 9 * Cin + SC
        wt = (torch.randn(N, kk, device=dev) / kk ** 0.5).to(BF); b = torch.randn(N, device=dev)
        kw = dict(bias=b, stride=stride, gn=gn_arg(gn, F_ * (H // stride) * (W // stride), N))
        if SC: kw["sc0"] = rnd(F_ * H * W, SC)
        return (lambda: ops.conv3x3(x, F_, H, W, wt, **kw)), N
    if op == "conv_t3":
        B, T_, HW, C, N, gn = key
        x = rnd(B * T_ * HW, C); wt = (torch.randn(N, 3 * C, device=dev) / (3 * C) ** 0.5).to(BF); b = torch.randn(N, device=dev)
        r = rnd(B * T_ * HW, N)
        g_ = gn_arg(gn, B * T_ * HW, N)
        return (lambda: ops.conv_t3(x, B, T_, HW, wt, bias=b, res1=r, gn=g_)), N
    if op == "upconv3x3":
        F_, H, W, C, N = key
        x = rnd(F_ * H * W, C); wp = ops.pack_upconv3x3(torch.randn(N, C, 3, 3, device=dev) / (9 * C) ** 0.5); b = torch.randn(N, device=dev)
        return (lambda: ops.upsample2x_conv3x3(x, F_, H, W, wp, bias=b)), N
rows = []
tot_auto = tot_best = 0.0
for (op, key), cnt in sorted(keys.items(), key=lambda kv: str(kv[0])):
    try:
        fn, N = make(op, key)
    except Exception as e:
        print("skip", op, key, e); continue
    ops.lib().ctrlv_igemm_override(0, 0, 0)
    base = t(fn); best = (base, "auto"); alt = {}
    for cg in (1, 2):
        for bn in (64, 96, 128, 160, 192, 256):
            if bn > N or (N % bn): continue
            ops.lib().ctrlv_igemm_override(bn, cg, 0)
            try: v = t(fn, 3)
            except Exception: v = float("nan")
            alt[f"cg{cg}/bn{bn}"] = v
            if v == v and v < best[0]: best = (v, f"cg{cg}/bn{bn}")
    ops.lib().ctrlv_igemm_override(0, 0, 0)
    tot_auto += base * cnt; tot_best += best[0] * cnt
    rows.append(dict(op=op, key=list(key), count=cnt, auto_us=base, best=best[1], best_us=best[0], alt=alt))
    print(f"{op:9s} {str(key):48s} n={cnt:3.0f} auto {base:7.1f}  best {best[1]:10s} {best[0]:7.1f}  gain/step {(base-best[0])*cnt:7.1f} us", flush=True)
print(f"TOTAL auto {tot_auto/1e3:.2f} ms  best {tot_best/1e3:.2f} ms")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_step_shapes.json", "w"), indent=1)
