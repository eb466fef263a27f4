"""Full-size temporal VAE decode (SVD widths, 14 x 320 x 512 output): new path vs the fp32 oracle on
the same GPU (PSNR / rel-L2), timings of both and of torch-eager bf16 of the oracle modules."""
import sys, os, json, math, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ctrlv_b200 import vae
from oracle import vae_oracle as V

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
T = int(os.environ.get("T", 14)); h = int(os.environ.get("H", 40)); w = int(os.environ.get("W", 64))
dev = "cuda"
torch.manual_seed(0)
ov = V.AutoencoderKLTemporalDecoder().to(dev).eval()
mv = vae.AutoencoderKLTemporalDecoder(state_dict=ov.state_dict())
g = torch.Generator("cpu").manual_seed(1234)
lat = torch.randn(1, T, 4, h, w, generator=g).to(dev)


def timed(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return out, e0.elapsed_time(e1) / n


got, ms_new = timed(lambda: vae.decode_latents(mv, lat, T, T))
with torch.no_grad():
    want, ms_f32 = timed(lambda: V.decode_latents(ov, lat, T, T), n=1)
    ob = ov.to(torch.bfloat16)
    ref_bf, ms_bf16 = timed(lambda: V.decode_latents(ob, lat.to(torch.bfloat16), T, T), n=2)


def psnr(a, b):
    mse = float(((a.float() - b.float()) ** 2).mean()); peak = float(b.max() - b.min())
    return 10 * math.log10(peak * peak / mse)


res = dict(shape=list(got.shape), ms_new=ms_new, ms_torch_fp32=ms_f32, ms_torch_bf16=ms_bf16,
           psnr_new_vs_fp32=psnr(got, want), rel_l2_new=float((got - want).norm() / want.norm()),
           psnr_torch_bf16_vs_fp32=psnr(ref_bf, want), rel_l2_torch_bf16=float((ref_bf.float() - want).norm() / want.norm()))
# encoder: 14 bbox frames at 320 x 512
x = (torch.rand(T, 3, 8 * h, 8 * w, generator=g) * 2 - 1).to(dev)
ov = ov.float()
gz, ms_enc = timed(lambda: mv.encode(x).latent_dist.mode())
with torch.no_grad():
    wz, ms_enc32 = timed(lambda: ov.encode_mode(x), n=1)
res.update(ms_encode_new=ms_enc, ms_encode_torch_fp32=ms_enc32, rel_l2_encode=float((gz - wz).norm() / wz.norm()))
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/vae_full.json", "w"), indent=1)
