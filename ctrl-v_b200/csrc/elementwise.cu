// Latency/HBM-bound glue of the denoise step: embeddings (sinusoid + small dense layers), the
// loop-body arithmetic (scale_model_input + concat, CFG combine + Euler-EDM update), nearest
// upsampling, residual adds and NCHW <-> channels-last conversion.
#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

// ------------------------------------------------------------------------------------------
// y[M][N] = act_out(act_in(x)[M][K] . W[N][K]^T + b): one warp per output column, rows in tiles of 4
// ------------------------------------------------------------------------------------------
__global__ void small_linear_kernel(const float* __restrict__ x, int M, int K,
                                    const bf16* __restrict__ W, const float* __restrict__ bias,
                                    int N, int act_in, int act_out, int accumulate,
                                    float* __restrict__ y) {
  pdl_wait();
  pdl_trigger();
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const bf16* wr = W + (size_t)n * K;
  for (int m0 = 0; m0 < M; m0 += 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane * 8; k < K; k += 256) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + k));
      float w[8];
      float2 f;
      f = unpack_bf16x2(u.x); w[0] = f.x; w[1] = f.y;
      f = unpack_bf16x2(u.y); w[2] = f.x; w[3] = f.y;
      f = unpack_bf16x2(u.z); w[4] = f.x; w[5] = f.y;
      f = unpack_bf16x2(u.w); w[6] = f.x; w[7] = f.y;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        if (m0 + mi < M) {
          const float* xr = x + (size_t)(m0 + mi) * K + k;
          const float4 a = __ldg(reinterpret_cast<const float4*>(xr));
          const float4 b = __ldg(reinterpret_cast<const float4*>(xr + 4));
          float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float t = act_in ? silu_acc_f(xv[j]) : xv[j];
            acc[mi] += t * w[j];
          }
        }
      }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const float s = warp_sum(acc[mi]);
      if (lane == 0 && m0 + mi < M) {
        float t = s + (bias ? bias[n] : 0.f);
        t = act_out ? silu_acc_f(t) : t;
        float* yp = y + (size_t)(m0 + mi) * N + n;
        *yp = accumulate ? *yp + t : t;
      }
    }
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0)
__global__ void sinusoid_kernel(const float* __restrict__ t, int n, int dim, int round_bf16,
                                float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim >> 1;
  if (idx >= n * half) return;
  const int i = idx / half, k = idx % half;
  // exp(-ln(10000) * k / half) in fp32, like torch
  const float freq = expf(-9.210340371976184f * (float)k / (float)half);
  const float a = t[i] * freq;
  float c = cosf(a), s = sinf(a);
  if (round_bf16) {
    c = __bfloat162float(__float2bfloat16(c));
    s = __bfloat162float(__float2bfloat16(s));
  }
  out[(size_t)i * dim + k] = c;
  out[(size_t)i * dim + half + k] = s;
}

// out[(bb*T+t)*hw + p][64] = [latents[b]/sqrt(sigma^2+1) (4) | image_latents[bb] (4) |
//                              control_cond[bb] (4) | zeros]
__global__ void prep_input_kernel(const float* __restrict__ lat, const float* __restrict__ img,
                                  const float* __restrict__ ctl, int B, int nb, int T, int hw,
                                  const float* __restrict__ sigma_dev, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const float sg = __ldg(sigma_dev);
  const float inv_scale = rsqrtf(sg * sg + 1.0f);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)nb * T * hw;
  if (idx >= total) return;
  const int p = (int)(idx % hw);
  const long long ft = idx / hw;
  const int t = (int)(ft % T);
  const int bb = (int)(ft / T);
  const int b = bb % B;
  float v[12];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    v[c] = lat[(((size_t)b * T + t) * 4 + c) * hw + p] * inv_scale;
    v[4 + c] = img ? img[(((size_t)bb * T + t) * 4 + c) * hw + p] : 0.f;
    v[8 + c] = ctl ? ctl[(((size_t)bb * T + t) * 4 + c) * hw + p] : 0.f;
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)idx * 64);
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  o[0] = u;
  u.x = pack_bf16x2(v[8], v[9]); u.y = pack_bf16x2(v[10], v[11]); u.z = 0; u.w = 0;
  o[1] = u;
  const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
  for (int i = 2; i < 8; ++i) o[i] = z;
}

// CFG combine + Euler (v-prediction) update, in fp32
__global__ void cfg_euler_kernel(float* __restrict__ lat, const float* __restrict__ noise, int ldn,
                                 int B, int cfg, int T, int hw, const float* __restrict__ guidance,
                                 const float* __restrict__ sigma_dev, int round_bf16) {
  pdl_wait();
  pdl_trigger();
  const float sigma = __ldg(sigma_dev), sigma_next = __ldg(sigma_dev + 1);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * hw;
  if (idx >= total) return;
  const int p = (int)(idx % hw);
  const long long ft = idx / hw;
  const int t = (int)(ft % T);
  const int b = (int)(ft / T);
  const float g = guidance ? guidance[t] : 1.f;
  const float c_out = -sigma * rsqrtf(sigma * sigma + 1.f);
  const float c_skip = 1.f / (sigma * sigma + 1.f);
  const float dt = sigma_next - sigma;
  const float* nu = noise + (((size_t)b * T + t) * hw + p) * ldn;
  const float* nc = noise + ((((size_t)B + b) * T + t) * hw + p) * ldn;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float v = nu[c];
    if (cfg) v = v + g * (nc[c] - v);
    float* xp = lat + (((size_t)b * T + t) * 4 + c) * hw + p;
    const float x = *xp;
    const float x0 = v * c_out + x * c_skip;
    const float d = (x - x0) / sigma;
    float xn = x + d * dt;
    if (round_bf16) xn = __bfloat162float(__float2bfloat16(xn));
    *xp = xn;
  }
}

__global__ void upsample2x_kernel(const uint4* __restrict__ src, int frames, int H, int W, int vpr,
                                  uint4* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)frames * 2 * H * 2 * W * vpr;
  if (idx >= total) return;
  const int v = (int)(idx % vpr);
  long long r = idx / vpr;
  const int ox = (int)(r % (2 * W)); r /= 2 * W;
  const int oy = (int)(r % (2 * H));
  const int f = (int)(r / (2 * H));
  out[idx] = __ldg(src + (((size_t)f * H + (oy >> 1)) * W + (ox >> 1)) * vpr + v);
}

__global__ void axpby_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, float a,
                             float b, long long nvec, uint4* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nvec) return;
  const uint4 ux = __ldg(x + idx), uy = __ldg(y + idx);
  uint4 o;
  float2 fx, fy;
  fx = unpack_bf16x2(ux.x); fy = unpack_bf16x2(uy.x); o.x = pack_bf16x2(a * fx.x + b * fy.x, a * fx.y + b * fy.y);
  fx = unpack_bf16x2(ux.y); fy = unpack_bf16x2(uy.y); o.y = pack_bf16x2(a * fx.x + b * fy.x, a * fx.y + b * fy.y);
  fx = unpack_bf16x2(ux.z); fy = unpack_bf16x2(uy.z); o.z = pack_bf16x2(a * fx.x + b * fy.x, a * fx.y + b * fy.y);
  fx = unpack_bf16x2(ux.w); fy = unpack_bf16x2(uy.w); o.w = pack_bf16x2(a * fx.x + b * fy.x, a * fx.y + b * fy.y);
  out[idx] = o;
}

// [frames][C][HW] (fp32 or bf16) -> channels c_off..c_off+C of out[frames*HW][Cpad] (bf16)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const T* __restrict__ src, int frames, int C, int HW, int Cpad,
                                    int c_off, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)frames * HW * C;
  if (idx >= total) return;
  // idx enumerates (f, c, p) with p fastest: coalesced reads; writes are strided but tiny (C <= 16)
  const int p = (int)(idx % HW);
  const long long r = idx / HW;
  const int c = (int)(r % C);
  const int f = (int)(r / C);
  out[((size_t)f * HW + p) * Cpad + c_off + c] = __float2bfloat16((float)src[idx]);
}

template <typename TI, typename TO>
__global__ void nhwc_to_nchw_kernel(const TI* __restrict__ src, long long ld, int frames, int C,
                                    int HW, TO* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)frames * HW * C;
  if (idx >= total) return;
  const int p = (int)(idx % HW);
  const long long r = idx / HW;
  const int c = (int)(r % C);
  const int f = (int)(r / C);
  out[idx] = (TO)(float)src[((size_t)f * HW + p) * ld + c];
}

}  // namespace ctrlv

using namespace ctrlv;

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

extern "C" int ctrlv_small_linear(const float* x, int32_t M, int32_t K, const void* W,
                                  const float* bias, int32_t N, int32_t act_in, int32_t act_out,
                                  int32_t accumulate, float* y, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(x && W && y, "small_linear: null pointer");
  CTRLV_CHECK_ARG(M > 0 && M <= 64 && K % 8 == 0 && N > 0, "small_linear: M=%d K=%d N=%d unsupported", M, K, N);
  const int wpb = 8;
  CTRLV_CUDA(launch_pdl(small_linear_kernel, dim3(nblk(N, wpb)), dim3(wpb * 32), (size_t)(0), stream, x, M, K, reinterpret_cast<const bf16*>(W),
                                                            bias, N, act_in, act_out, accumulate, y));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_sinusoid(const float* t, int32_t n, int32_t dim, int32_t round_bf16, float* out,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(t && out && n > 0 && dim > 0 && dim % 2 == 0, "sinusoid: bad arguments");
  CTRLV_CUDA(launch_pdl(sinusoid_kernel, dim3(nblk((long long)n * dim / 2, 128)), dim3(128), (size_t)(0), stream, t, n, dim, round_bf16, out));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_prep_input(const float* latents, const float* image_latents,
                                const float* control_cond, int32_t B, int32_t cfg, int32_t T,
                                int32_t h, int32_t w, const float* sigma_dev, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(latents && out && sigma_dev, "prep_input: null pointer");
  const int nb = cfg ? 2 * B : B;
  const long long total = (long long)nb * T * h * w;
  CTRLV_CUDA(launch_pdl(prep_input_kernel, dim3(nblk(total, 256)), dim3(256), (size_t)(0), stream, latents, image_latents, control_cond, B, nb,
                                                          T, h * w, sigma_dev, reinterpret_cast<bf16*>(out)));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_cfg_euler(float* latents, const float* noise, int32_t ld_noise, int32_t B,
                               int32_t cfg, int32_t T, int32_t h, int32_t w, const float* guidance,
                               const float* sigma_dev, int32_t round_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(latents && noise && sigma_dev && ld_noise >= 4, "cfg_euler: bad arguments");
  const long long total = (long long)B * T * h * w;
  CTRLV_CUDA(launch_pdl(cfg_euler_kernel, dim3(nblk(total, 256)), dim3(256), (size_t)(0), stream, latents, noise, ld_noise, B, cfg, T, h * w,
                                                         guidance, sigma_dev, round_bf16));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_upsample2x(const void* src, int32_t frames, int32_t H, int32_t Wd, int32_t C,
                                void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(src && out && C % 8 == 0, "upsample2x: bad arguments");
  const long long total = (long long)frames * 4 * H * Wd * (C / 8);
  CTRLV_CUDA(launch_pdl(upsample2x_kernel, dim3(nblk(total, 256)), dim3(256), (size_t)(0), stream, reinterpret_cast<const uint4*>(src), frames, H,
                                                          Wd, C / 8, reinterpret_cast<uint4*>(out)));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_axpby(const void* x, const void* y, float a, float b, int64_t n, void* out,
                           void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(x && y && out && n % 8 == 0, "axpby: bad arguments (n %% 8)");
  CTRLV_CUDA(launch_pdl(axpby_kernel, dim3(nblk(n / 8, 256)), dim3(256), (size_t)(0), stream, reinterpret_cast<const uint4*>(x),
                                                     reinterpret_cast<const uint4*>(y), a, b, n / 8,
                                                     reinterpret_cast<uint4*>(out)));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_nchw_to_nhwc(const void* src, int32_t src_is_f32, int32_t frames, int32_t C,
                                  int32_t HW, int32_t Cpad, int32_t c_off, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(src && out && c_off >= 0 && c_off + C <= Cpad, "nchw_to_nhwc: bad arguments");
  const long long total = (long long)frames * HW * C;
  if (src_is_f32)
    CTRLV_CUDA(launch_pdl(nchw_to_nhwc_kernel<float>, dim3(nblk(total, 256)), dim3(256), (size_t)(0), stream, 
        reinterpret_cast<const float*>(src), frames, C, HW, Cpad, c_off, reinterpret_cast<bf16*>(out)));
  else
    CTRLV_CUDA(launch_pdl(nchw_to_nhwc_kernel<bf16>, dim3(nblk(total, 256)), dim3(256), (size_t)(0), stream, 
        reinterpret_cast<const bf16*>(src), frames, C, HW, Cpad, c_off, reinterpret_cast<bf16*>(out)));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}

extern "C" int ctrlv_nhwc_to_nchw(const void* src, int32_t src_is_f32, int64_t ld, int32_t frames,
                                  int32_t C, int32_t HW, int32_t out_is_f32, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(src && out, "nhwc_to_nchw: null pointer");
  const long long total = (long long)frames * HW * C;
  const unsigned g = nblk(total, 256);
  if (src_is_f32 && out_is_f32)
    CTRLV_CUDA(launch_pdl(nhwc_to_nchw_kernel<float, float>, dim3(g), dim3(256), (size_t)(0), stream, reinterpret_cast<const float*>(src), ld, frames, C, HW, reinterpret_cast<float*>(out)));
  else if (src_is_f32)
    CTRLV_CUDA(launch_pdl(nhwc_to_nchw_kernel<float, bf16>, dim3(g), dim3(256), (size_t)(0), stream, reinterpret_cast<const float*>(src), ld, frames, C, HW, reinterpret_cast<bf16*>(out)));
  else if (out_is_f32)
    CTRLV_CUDA(launch_pdl(nhwc_to_nchw_kernel<bf16, float>, dim3(g), dim3(256), (size_t)(0), stream, reinterpret_cast<const bf16*>(src), ld, frames, C, HW, reinterpret_cast<float*>(out)));
  else
    CTRLV_CUDA(launch_pdl(nhwc_to_nchw_kernel<bf16, bf16>, dim3(g), dim3(256), (size_t)(0), stream, reinterpret_cast<const bf16*>(src), ld, frames, C, HW, reinterpret_cast<bf16*>(out)));
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}
