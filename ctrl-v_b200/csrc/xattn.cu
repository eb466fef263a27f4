// Cross-attention over a short context (diffusers Attention with encoder_hidden_states [batch, L, D], L > 1,
// as BasicTransformerBlock.attn2 / TemporalBasicTransformerBlock.attn2 receive it through
// controlnet.py:230,244-245).  The Box2Video pipelines only ever produce L = 1 (handled without a kernel:
// softmax over one key is 1, see models._CrossAttnL1); this kernel closes the signature for general L <= 256.
// CUDA cores: the work is 4*L*64 FLOP per (row, head) — three orders of magnitude below the self-attention.
//
// One warp per (query row, head): phase 1 lanes = keys (scores, softmax), phase 2 lanes = channel pairs (P V).
#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

constexpr int kXaMaxL = 256;

__device__ __forceinline__ int xa_ctx_index(int mode, int m, int div, int mod, int nB) {
  const int a = m / div;
  if (mode == 1) return a;
  if (mode == 2) return a % mod;
  return (a * mod + m % mod) % nB;
}

__global__ void __launch_bounds__(256) cross_attn_kernel(const bf16* __restrict__ q, long long ldq,
                                                         const bf16* __restrict__ kv, long long ldkv, int C, int M,
                                                         int heads, int L, float scale_log2e, int mode, int div, int mod,
                                                         int nB, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sq[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long unit = (long long)blockIdx.x * 8 + warp;
  if (unit >= (long long)M * heads) return;
  const int m = (int)(unit / heads), h = (int)(unit % heads);
  const int ctx = xa_ctx_index(mode, m, div, mod, nB);
  const float2 q2 = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(q + (size_t)m * ldq + h * 64 + 2 * lane));
  sq[warp][2 * lane] = q2.x * scale_log2e;
  sq[warp][2 * lane + 1] = q2.y * scale_log2e;
  __syncwarp();
  const bf16* kbase = kv + (size_t)ctx * L * ldkv + h * 64;
  const bf16* vbase = kbase + C;
  float s[kXaMaxL / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int r = 0; r < kXaMaxL / 32; ++r) {
    const int l = r * 32 + lane;
    s[r] = -INFINITY;
    if (l < L) {
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)l * ldkv);
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 u = __ldg(kr + c);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          acc = fmaf(sq[warp][c * 8 + 2 * k], f.x, acc);
          acc = fmaf(sq[warp][c * 8 + 2 * k + 1], f.y, acc);
        }
      }
      s[r] = acc;
      mx = fmaxf(mx, acc);
    }
  }
  mx = warp_max(mx);
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < kXaMaxL / 32; ++r) {
    s[r] = (r * 32 + lane < L) ? ex2_approx(s[r] - mx) : 0.f;
    sum += s[r];
  }
  sum = warp_sum(sum);
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int r = 0; r < kXaMaxL / 32; ++r) {
    if (r * 32 < L) {  // warp-uniform
      const int n = min(32, L - r * 32);
      for (int j = 0; j < n; ++j) {
        const float pj = __shfl_sync(0xffffffffu, s[r], j);
        const float2 v2 = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(vbase + (size_t)(r * 32 + j) * ldkv + 2 * lane)));
        a0 = fmaf(pj, v2.x, a0);
        a1 = fmaf(pj, v2.y, a1);
      }
    }
  }
  const float inv = 1.0f / sum;
  *reinterpret_cast<uint32_t*>(out + (size_t)m * C + h * 64 + 2 * lane) = pack_bf16x2(a0 * inv, a1 * inv);
}

}  // namespace ctrlv

using namespace ctrlv;

extern "C" int ctrlv_cross_attn(const void* q, int64_t ldq, const void* kv, int64_t ldkv, int32_t M, int32_t heads,
                                int32_t L, int32_t n_ctx, float scale, int32_t ctx_mode, int32_t ctx_div,
                                int32_t ctx_mod, int32_t ctx_B, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(q && kv && out && M > 0 && heads > 0, "cross_attn: bad arguments");
  CTRLV_CHECK_ARG(L >= 1 && L <= kXaMaxL, "cross_attn: context length %d outside [1, %d]", L, kXaMaxL);
  CTRLV_CHECK_ARG(ldq % 2 == 0 && ldkv % 8 == 0 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(q) & 3) == 0, "cross_attn: q needs 4-byte, kv 16-byte aligned rows");
  CTRLV_CHECK_ARG(ctx_mode >= 1 && ctx_mode <= 3 && ctx_div > 0 && (ctx_mode == 1 || ctx_mod > 0) && (ctx_mode != 3 || ctx_B > 0),
                  "cross_attn: bad context index mode");
  // the largest context index any row can produce must exist
  const int last = ctx_mode == 1 ? (M - 1) / ctx_div : (ctx_mode == 2 ? ctx_mod - 1 : ctx_B - 1);
  CTRLV_CHECK_ARG(last < n_ctx, "cross_attn: rows index context %d but only %d contexts were given", last, n_ctx);
  const int C = heads * 64;
  const long long units = (long long)M * heads;
  const unsigned blocks = (unsigned)((units + 7) / 8);
  CTRLV_CUDA(launch_pdl(cross_attn_kernel, dim3(blocks), dim3(256), (size_t)0, stream, reinterpret_cast<const bf16*>(q),
                        (long long)ldq, reinterpret_cast<const bf16*>(kv), (long long)ldkv, C, M, heads, L,
                        scale * 1.4426950408889634f, ctx_mode, ctx_div, ctx_mod, ctx_B, reinterpret_cast<bf16*>(out)));
  return CTRLV_OK;
}
