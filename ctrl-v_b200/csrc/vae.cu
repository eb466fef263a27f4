// Glue kernels of the temporal VAE (diffusers AutoencoderKLTemporalDecoder):
//  * row softmax for the single-head, 512-wide mid-block attention (scores come from the igemm as
//    fp32; probabilities go back as the bf16 A operand of the P.V GEMM);
//  * time_conv_out: the (3,1,1) Conv3d over 3 image channels fused with the channels-last ->
//    NCHW conversion of the decoder output.
#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

// one warp per row; the row is read twice from L2 (max+sum online in one pass, then write)
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long ld_s, int M, int N,
                                    float scale_log2e, bf16* __restrict__ p, long long ld_p) {
  pdl_wait();
  pdl_trigger();
  const int warp = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* row = s + (size_t)warp * ld_s;
  float m = -INFINITY, l = 0.f;
  for (int i = lane * 4; i < N; i += 128) {
    float v[4];
    if (i + 4 <= N && (ld_s & 3) == 0) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(row + i));
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (i + j < N) ? __ldg(row + i + j) : -INFINITY;
    }
    const float bm = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    if (bm > m) {
      l *= ex2_approx((m - bm) * scale_log2e);
      m = bm;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) l += ex2_approx((v[j] - m) * scale_log2e);
  }
  // combine the lanes' (max, sum) pairs
  float mm = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
  l = (m == -INFINITY) ? 0.f : l * ex2_approx((m - mm) * scale_log2e);
  l = warp_sum(l);
  const float inv = 1.0f / l;
  bf16* out = p + (size_t)warp * ld_p;
  for (int i = lane * 4; i < N; i += 128) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (i + j < N)
        out[i + j] = __float2bfloat16(ex2_approx((__ldg(row + i + j) - mm) * scale_log2e) * inv);
  }
}

// x: [B][T][HW][ld] fp32 channels-last (first C columns), w: [C][C][3] (Conv3d weight [co][ci][kt][1][1]),
// out: [B*T][C][HW] fp32.  Zero padding along T per clip.
__global__ void time_conv_out_kernel(const float* __restrict__ x, int ld, int B, int T, int HW, int C,
                                     const float* __restrict__ w, const float* __restrict__ bias,
                                     float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)B * T * HW;
  if (idx >= total) return;
  const int p = (int)(idx % HW);
  const int t = (int)((idx / HW) % T);
  const int b = (int)(idx / ((long long)HW * T));
  float acc[4];
#pragma unroll
  for (int co = 0; co < 4; ++co) acc[co] = co < C ? __ldg(bias + co) : 0.f;
#pragma unroll
  for (int kt = 0; kt < 3; ++kt) {
    const int tt = t + kt - 1;
    if (tt < 0 || tt >= T) continue;
    const float* xi = x + (((size_t)b * T + tt) * HW + p) * ld;
    for (int ci = 0; ci < C; ++ci) {
      const float v = __ldg(xi + ci);
#pragma unroll
      for (int co = 0; co < 4; ++co)
        if (co < C) acc[co] = fmaf(__ldg(w + (co * C + ci) * 3 + kt), v, acc[co]);
    }
  }
  for (int co = 0; co < C; ++co) out[(((size_t)b * T + t) * C + co) * HW + p] = acc[co];
}

}  // namespace ctrlv

using namespace ctrlv;

extern "C" int ctrlv_softmax_rows(const float* scores, int64_t ld_scores, int32_t M, int32_t N, float scale,
                                  void* probs, int64_t ld_probs, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(scores && probs && M > 0 && N > 0, "softmax_rows: bad arguments");
  const long long threads = (long long)M * 32;
  CTRLV_CUDA(launch_pdl(softmax_rows_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), (size_t)0, stream,
                        scores, (long long)ld_scores, M, N, scale * 1.4426950408889634f,
                        reinterpret_cast<bf16*>(probs), (long long)ld_probs));
  return CTRLV_OK;
}

extern "C" int ctrlv_time_conv_out(const float* x, int32_t ld, int32_t B, int32_t T, int32_t HW, int32_t C,
                                   const float* w, const float* bias, float* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(x && w && bias && out && B > 0 && T > 0 && HW > 0, "time_conv_out: bad arguments");
  CTRLV_CHECK_ARG(C >= 1 && C <= 4 && ld >= C, "time_conv_out: C=%d unsupported (1..4)", C);
  const long long total = (long long)B * T * HW;
  CTRLV_CUDA(launch_pdl(time_conv_out_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)0, stream, x,
                        ld, B, T, HW, C, w, bias, out));
  return CTRLV_OK;
}
