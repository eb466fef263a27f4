// Image-conditioning prologue of the pipelines: the antialiased resize of the
// conditioning image to the CLIP resolution (Gaussian blur with reflect padding, then bicubic
// interpolation with align_corners=True) and the patch gather that turns the CLIP patch-embedding
// convolution (kernel = stride = patch) into a GEMM.  Small fp32 images; one thread per output.
#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

__device__ __forceinline__ int reflect_idx(int i, int n) {  // F.pad(mode="reflect"): no edge repeat
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}

// 1-D Gaussian along x (axis 0) or y (axis 1) of [planes][H][W]; taps[ks] normalised on the host;
// pad_front = (ks - 1) / 2 as in the reference's _compute_padding.
__global__ void blur1d_kernel(const float* __restrict__ src, int planes, int H, int W, int axis,
                              const float* __restrict__ taps, int ks, float* __restrict__ dst) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)planes * H * W;
  if (idx >= total) return;
  const int x = (int)(idx % W);
  const int y = (int)((idx / W) % H);
  const long long pl = idx / ((long long)W * H);
  const float* p = src + pl * H * W;
  const int front = (ks - 1) / 2;
  float acc = 0.f;
  if (axis == 0) {
    for (int k = 0; k < ks; ++k) acc = fmaf(__ldg(taps + k), __ldg(p + (size_t)y * W + reflect_idx(x + k - front, W)), acc);
  } else {
    for (int k = 0; k < ks; ++k) acc = fmaf(__ldg(taps + k), __ldg(p + (size_t)reflect_idx(y + k - front, H) * W + x), acc);
  }
  dst[idx] = acc;
}

__device__ __forceinline__ void cubic_coeffs(float t, float* w) {  // A = -0.75 (torch upsample_bicubic2d)
  const float A = -0.75f;
  float x = t + 1.0f;
  w[0] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
  x = t;
  w[1] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 1.0f - t;
  w[2] = ((A + 2.0f) * x - (A + 3.0f)) * x * x + 1.0f;
  x = 2.0f - t;
  w[3] = ((A * x - 5.0f * A) * x + 8.0f * A) * x - 4.0f * A;
}

// F.interpolate(mode="bicubic", align_corners=True): source coordinate = dst * (in-1)/(out-1),
// neighbours clamped to the image.
__global__ void bicubic_ac_kernel(const float* __restrict__ src, int planes, int H, int W, int Ho, int Wo,
                                  float* __restrict__ dst) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)planes * Ho * Wo;
  if (idx >= total) return;
  const int ox = (int)(idx % Wo);
  const int oy = (int)((idx / Wo) % Ho);
  const long long pl = idx / ((long long)Wo * Ho);
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const float fy = sy * oy, fx = sx * ox;
  const int iy = (int)floorf(fy), ix = (int)floorf(fx);
  float wy[4], wx[4];
  cubic_coeffs(fy - iy, wy);
  cubic_coeffs(fx - ix, wx);
  const float* p = src + pl * H * W;
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), H - 1);
    float row = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = min(max(ix - 1 + b, 0), W - 1);
      row = fmaf(wx[b], __ldg(p + (size_t)yy * W + xx), row);
    }
    acc = fmaf(wy[a], row, acc);
  }
  dst[idx] = acc;
}

// rows[(b*gh + py)*gw + px][c*P*P + iy*P + ix] = (u - mean[c]) / std[c] with u = a*img[b][c][..] + s
// (clamped to [0, 1] when clamp01 is set), zero-padded to Kpad columns (bf16): the A operand of the
// patch-embedding GEMM.
__global__ void clip_patchify_kernel(const float* __restrict__ img, int B, int C, int H, int W, int P,
                                     float a, float s, int clamp01, const float* __restrict__ mean,
                                     const float* __restrict__ stdv, int Kpad, bf16* __restrict__ rows) {
  pdl_wait();
  pdl_trigger();
  const int gh = H / P, gw = W / P;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * gh * gw * Kpad;
  if (idx >= total) return;
  const int k = (int)(idx % Kpad);
  const long long r = idx / Kpad;
  float v = 0.f;
  if (k < C * P * P) {
    const int ix = k % P, iy = (k / P) % P, c = k / (P * P);
    const int px = (int)(r % gw), py = (int)((r / gw) % gh);
    const long long b = r / ((long long)gw * gh);
    const float x = __ldg(img + ((b * C + c) * H + (size_t)py * P + iy) * W + (size_t)px * P + ix);
    float u = a * x + s;
    if (clamp01) u = fminf(fmaxf(u, 0.f), 1.f);
    v = (u - __ldg(mean + c)) / __ldg(stdv + c);
  }
  rows[idx] = __float2bfloat16(v);
}

}  // namespace ctrlv

using namespace ctrlv;

static inline unsigned nblk256(long long n) { return (unsigned)((n + 255) / 256); }

extern "C" int ctrlv_blur1d_reflect(const float* src, int32_t planes, int32_t H, int32_t W, int32_t axis,
                                    const float* taps, int32_t ks, float* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(src && dst && taps && src != dst, "blur1d_reflect: bad pointers");
  CTRLV_CHECK_ARG(planes > 0 && H > 0 && W > 0 && ks >= 1 && (axis == 0 || axis == 1), "blur1d_reflect: bad shape");
  CTRLV_CHECK_ARG((ks - 1) / 2 < (axis == 0 ? W : H) && ks / 2 < (axis == 0 ? W : H),
                  "blur1d_reflect: reflect padding %d needs a larger image", ks / 2);
  CTRLV_CUDA(launch_pdl(blur1d_kernel, dim3(nblk256((long long)planes * H * W)), dim3(256), (size_t)0, stream, src,
                        planes, H, W, axis, taps, ks, dst));
  return CTRLV_OK;
}

extern "C" int ctrlv_resize_bicubic_ac(const float* src, int32_t planes, int32_t H, int32_t W, int32_t Ho,
                                       int32_t Wo, float* dst, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(src && dst && planes > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0, "resize_bicubic_ac: bad arguments");
  CTRLV_CUDA(launch_pdl(bicubic_ac_kernel, dim3(nblk256((long long)planes * Ho * Wo)), dim3(256), (size_t)0, stream,
                        src, planes, H, W, Ho, Wo, dst));
  return CTRLV_OK;
}

extern "C" int ctrlv_clip_patchify(const float* img, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P,
                                   float a, float s, int32_t clamp01, const float* mean, const float* stdv,
                                   int32_t Kpad, void* rows, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(img && mean && stdv && rows, "clip_patchify: null pointer");
  CTRLV_CHECK_ARG(B > 0 && C > 0 && P > 0 && H % P == 0 && W % P == 0 && Kpad >= C * P * P && Kpad % 8 == 0,
                  "clip_patchify: bad shape");
  const long long total = (long long)B * (H / P) * (W / P) * Kpad;
  CTRLV_CUDA(launch_pdl(clip_patchify_kernel, dim3(nblk256(total)), dim3(256), (size_t)0, stream, img, B, C, H, W, P,
                        a, s, clamp01, mean, stdv, Kpad, reinterpret_cast<bf16*>(rows)));
  return CTRLV_OK;
}
