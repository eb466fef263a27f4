// tcgen05 flash attention, head_dim 64, no mask — the core of diffusers' AttnProcessor2_0
// (F.scaled_dot_product_attention) for both attention flavours of TransformerSpatioTemporalModel.
//
//  attn_kernel (spatial, short sequences): one CTA per (128-query tile, head, frame); keys/values
//                    streamed in blocks of 128 through a 3-slot TMA ring; online softmax.
//  attn2_kernel (spatial, S >= 160): two query tiles per CTA, P in TMEM (below).
//  tattn_kernel (temporal): sequences are the T frames of one spatial site.  G = 128/T sites are packed
//                    into one 128-row tile (row = t*G + g) fetched by ONE 4-D TMA box straight
//                    from the [B][T][S][3C] projection buffer (no permute copies); a block-diagonal
//                    mask (same site) restricts each query to its own T keys.  HBM-bound (below).
//
//  S = Q K^T      : tcgen05.mma  M=128 N=128 K=64, both operands K-major from TMA (128B swizzle)
//  softmax        : one thread per query row (TMEM lane == row): no shuffles; exp2 on raw scores
//  O_blk = P V    : P written as bf16 to swizzled smem (K-major A operand), V used as loaded
//                   (MN-major B operand); block result folded into fp32 registers with the
//                   running-max rescale.
//  TMEM: S 128 cols + O 64 cols (256 allocated) and ~97 KB smem -> 2 CTAs per SM overlap each
//  other's softmax and MMA phases.
#include <cstdlib>

#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

constexpr int kAttnThreads = 64 + 256;  // TMA warp, MMA warp, 8 softmax warps
constexpr int kTile = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16

struct AttnParams {
  CUtensorMap tm;
  int C, S, T, G, heads, nkv, box_bytes;
  float c;  // softmax scale * log2(e)
  bf16* out;
};

__device__ __forceinline__ float max3f(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2) attn_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, s_full, p_full, o_full;
  __shared__ __align__(8) uint64_t kv_full[3], kv_empty[3];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xmax[2][2][128];  // [block parity][column half][row]: partial row max exchange
  __shared__ float xsum[2][128];     // [column half][row]: partial row sums (end of the loop)

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + kTile;      // 3 slots
  uint8_t* sP = smem + 4 * kTile;   // 2 atoms of [128 x 64]
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int nkv = p.nkv;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    mbar_init(&q_full, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 8);
    mbar_init(&o_full, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  const uint32_t tS = tmem_base;
  const uint32_t tO = tmem_base + 128;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const int colq = head * 64, colk = p.C + head * 64, colv = 2 * p.C + head * 64;
      mbar_expect_tx(&q_full, (uint32_t)p.box_bytes);
      tma_load_3d(sQ, &p.tm, &q_full, colq, blockIdx.x * 128, blockIdx.z);
      for (int n = 0; n < 2 * nkv; ++n) {
        const int slot = n % 3;
        const uint32_t ph = (uint32_t)((n / 3) & 1);
        mbar_wait(&kv_empty[slot], ph ^ 1);
        mbar_expect_tx(&kv_full[slot], (uint32_t)p.box_bytes);
        const int j = n >> 1;
        const int col = (n & 1) ? colv : colk;
        tma_load_3d(sKV + slot * kTile, &p.tm, &kv_full[slot], col, j * 128, blockIdx.z);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aKV = smem_u32(sKV);
      auto issue_qk = [&](int j) {
        const int n = 2 * j, slot = n % 3;
        mbar_wait(&kv_full[slot], (uint32_t)((n / 3) & 1));
        tc_fence_after();
        const uint32_t aK = aKV + slot * kTile;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tS, make_sdesc(aQ + k * 32, 16, 1024), make_sdesc(aK + k * 32, 16, 1024), idesc_qk,
                  (uint32_t)(k != 0));
        umma_commit(&kv_empty[slot]);
        umma_commit(&s_full);
      };
      mbar_wait(&q_full, 0);
      issue_qk(0);
      for (int j = 0; j < nkv; ++j) {
        mbar_wait(&p_full, (uint32_t)(j & 1));  // P_j in smem, S_j consumed
        tc_fence_after();
        if (j + 1 < nkv) issue_qk(j + 1);
        const int n = 2 * j + 1, slot = n % 3;
        mbar_wait(&kv_full[slot], (uint32_t)((n / 3) & 1));
        tc_fence_after();
        const uint32_t aV = aKV + slot * kTile;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ss(tO, make_sdesc(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                  make_sdesc(aV + k * 2048, 1024, 1024), idesc_pv, (uint32_t)(k != 0));
        umma_commit(&kv_empty[slot]);
        umma_commit(&o_full);
      }
    }
  } else {
    // ============================ softmax / epilogue ==============================
    // 8 warps: two per TMEM lane quarter.  Warp (q, half) owns key columns [64*half, 64*half+64) of
    // its 32 query rows and output columns [32*half, 32*half+32); the two halves exchange their
    // partial row max through shared memory (the partial row sums only meet at the end).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    auto chunk_mask = [&](int kv0, int c) -> uint32_t {  // valid key columns of chunk c (ragged last block)
      const int nvalid = p.S - kv0 - c * 32;
      return nvalid >= 32 ? 0xffffffffu : (nvalid <= 0 ? 0u : ((1u << nvalid) - 1u));
    };
    uint8_t* prow = sP + half * kTile + r * 128;  // this warp's 64 key columns = one 128B-swizzle atom

    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full, (uint32_t)(j & 1));
      tc_fence_after();
      const int kv0 = j * 128;
      // ---- pass 1: row max over this warp's valid columns
      float m_blk = -INFINITY;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t raw[32];
        tmem_ld32(tS + lane_off + c * 32, raw);
        tmem_ld_wait();
        const uint32_t okm = chunk_mask(kv0, c);
        if (okm == 0xffffffffu) {  // common case: every column valid, no per-element predicate
          float m0 = fmaxf(__uint_as_float(raw[0]), __uint_as_float(raw[1]));
          float m1 = fmaxf(__uint_as_float(raw[2]), __uint_as_float(raw[3]));
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            m0 = max3f(m0, __uint_as_float(raw[i]), __uint_as_float(raw[i + 1]));
            m1 = max3f(m1, __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
          }
          m_blk = max3f(m_blk, m0, m1);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if ((okm >> i) & 1u) m_blk = fmaxf(m_blk, __uint_as_float(raw[i]));
        }
      }
      // exchange with the other half (double-buffered by block parity, one 256-thread barrier)
      xmax[j & 1][half][r] = m_blk;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      m_blk = fmaxf(m_blk, xmax[j & 1][half ^ 1][r]);
      const float m_new = fmaxf(m_run, m_blk);
      const float alpha = ex2f((m_run - m_new) * p.c);
      // ---- fold the previous block's P V into the fp32 accumulator (also frees the P buffer)
      if (j > 0) {
        mbar_wait(&o_full, (uint32_t)((j - 1) & 1));
        tc_fence_after();
        uint32_t raw[32];
        tmem_ld32(tO + lane_off + half * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[i] += __uint_as_float(raw[i]);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] *= alpha;
      l_run *= alpha;
      m_run = m_new;
      const float mc = m_new * p.c;
      // ---- pass 2: P = exp2(s*c - m*c) as bf16 into swizzled smem
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;
        uint32_t raw[32];
        tmem_ld32(tS + lane_off + c * 32, raw);
        tmem_ld_wait();
        uint32_t pk[16];
        const uint32_t okm = chunk_mask(kv0, c);
        if (okm == 0xffffffffu) {
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ex2f(fmaf(__uint_as_float(raw[i]), p.c, -mc));
            const float p1 = ex2f(fmaf(__uint_as_float(raw[i + 1]), p.c, -mc));
            l0 += p0; l1 += p1;
            pk[i >> 1] = pack_bf16x2(p0, p1);
          }
          l_run += l0 + l1;
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float pv[2];
#pragma unroll
            for (int e = 0; e < 2; ++e)
              pv[e] = ((okm >> (i + e)) & 1u) ? ex2f(__uint_as_float(raw[i + e]) * p.c - mc) : 0.f;
            l_run += pv[0] + pv[1];
            pk[i >> 1] = pack_bf16x2(pv[0], pv[1]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = cc * 4 + i;
          *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full);
    }
    // ---- last block's P V, combine the two partial row sums, normalise, store
    mbar_wait(&o_full, (uint32_t)((nkv - 1) & 1));
    tc_fence_after();
    {
      uint32_t raw[32];
      tmem_ld32(tO + lane_off + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[i] += __uint_as_float(raw[i]);
    }
    xsum[half][r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float inv = 1.0f / (l_run + xsum[half ^ 1][r]);
    const int s = blockIdx.x * 128 + r;
    const bool valid = s < p.S;
    const long long orow = (long long)blockIdx.z * p.S + s;
    if (valid) {
      bf16* op = p.out + orow * p.C + head * 64 + half * 32;  // 64 B = two whole sectors, 256-bit stores
#pragma unroll
      for (int i = 0; i < 2; ++i)
        st_global_v8(op + 16 * i,
                     pack_bf16x2(o_acc[16 * i] * inv, o_acc[16 * i + 1] * inv), pack_bf16x2(o_acc[16 * i + 2] * inv, o_acc[16 * i + 3] * inv),
                     pack_bf16x2(o_acc[16 * i + 4] * inv, o_acc[16 * i + 5] * inv), pack_bf16x2(o_acc[16 * i + 6] * inv, o_acc[16 * i + 7] * inv),
                     pack_bf16x2(o_acc[16 * i + 8] * inv, o_acc[16 * i + 9] * inv), pack_bf16x2(o_acc[16 * i + 10] * inv, o_acc[16 * i + 11] * inv),
                     pack_bf16x2(o_acc[16 * i + 12] * inv, o_acc[16 * i + 13] * inv), pack_bf16x2(o_acc[16 * i + 14] * inv, o_acc[16 * i + 15] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------
// Temporal attention (sequences = the T frames of one spatial site; one key block): the single-block special case
// of the one-tile kernel above (its former temporal mode), shrunk so that FOUR CTAs share an SM instead of two.  The launch is a stream of short
// dependent chains (TMA Q|K|V -> S = Q K^T -> masked softmax -> P -> O = P V -> store, about 6 us per CTA) with
// almost no arithmetic (HBM-bound: 8 * rows * C bytes), so its speed is the number of chains in flight per SM:
//   * 4 softmax warps (one per TMEM lane quarter, a thread owns a whole 128-column score row): no exchange between
//     column halves, 192 threads, no running (max, sum, O) state in registers;
//   * shared memory 48 KB: P (128 x 128 bf16) is written over Q | K once S has been computed;
//   * TMEM 128 columns: O (64) is accumulated over the first half of S after every softmax warp has read S;
//   * only the rows past G * T of the Q, K, V tiles are zeroed (the TMA box covers rows [0, G * T)).
// ------------------------------------------------------------------------------------------
constexpr int kTattnThreads = 64 + 128;  // TMA warp, MMA warp, 4 softmax warps
constexpr int kTattnSmem = 3 * kTile + 1024;

__global__ void __launch_bounds__(kTattnThreads, 4) tattn_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t qkv_full, s_full, p_full, o_full;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;
  uint8_t* sV = smem + 2 * kTile;
  uint8_t* sP = smem;  // two [128 x 64] atoms over Q | K (dead once S = Q K^T is complete)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  pdl_trigger();

  {  // rows the box does not cover must be finite: 0 * garbage would be NaN in P V (V) and in the masked scores
    const int row0 = p.G * p.T, nz = (128 - row0) * 8;  // 16-byte pieces per tile
    for (int i = threadIdx.x; i < 3 * nz; i += kTattnThreads) {
      const int tile = i / nz, rem = i - tile * nz;
      *reinterpret_cast<uint4*>(smem + (size_t)tile * kTile + (size_t)(row0 + (rem >> 3)) * 128 + (rem & 7) * 16) =
          make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    mbar_init(&qkv_full, 1);
    mbar_init(&s_full, 1);
    mbar_init(&p_full, 4);
    mbar_init(&o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tS = tmem_base_smem;
  const uint32_t tO = tS;  // O over S[:, 0:64)
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      const int colq = head * 64, colk = p.C + head * 64, colv = 2 * p.C + head * 64;
      mbar_expect_tx(&qkv_full, (uint32_t)(3 * p.box_bytes));
      tma_load_4d(sQ, &p.tm, &qkv_full, colq, blockIdx.x * p.G, 0, blockIdx.z);
      tma_load_4d(sK, &p.tm, &qkv_full, colk, blockIdx.x * p.G, 0, blockIdx.z);
      tma_load_4d(sV, &p.tm, &qkv_full, colv, blockIdx.x * p.G, 0, blockIdx.z);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc(128, 128, 0, 0);
      const uint32_t idesc_pv = make_idesc(128, 64, 0, 1);  // B (= V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_wait(&qkv_full, 0);
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_ss(tS, make_sdesc(aQ + k * 32, 16, 1024), make_sdesc(aK + k * 32, 16, 1024), idesc_qk, (uint32_t)(k != 0));
      umma_commit(&s_full);
      mbar_wait(&p_full, 0);  // P in shared memory, S read by every softmax warp
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_ss(tO, make_sdesc(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024), make_sdesc(aV + k * 2048, 1024, 1024),
                idesc_pv, (uint32_t)(k != 0));
      umma_commit(&o_full);
    }
  } else {
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const int r = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const int rg = r % p.G;
    uint32_t tmask[4] = {0u, 0u, 0u, 0u};  // the T keys of this row's site sit at columns rg + t * G
    for (int t = 0; t < p.T; ++t) {
      const int col = rg + t * p.G;
      tmask[col >> 5] |= 1u << (col & 31);
    }
    mbar_wait(&s_full, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t raw[32];
      tmem_ld32(tS + lane_off + c * 32, raw);
      tmem_ld_wait();
      const uint32_t okm = c == 0 ? tmask[0] : (c == 1 ? tmask[1] : (c == 2 ? tmask[2] : tmask[3]));
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((okm >> i) & 1u) m = fmaxf(m, __uint_as_float(raw[i]));
    }
    const float mc = m * p.c;
    float l = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t raw[32];
      tmem_ld32(tS + lane_off + c * 32, raw);
      tmem_ld_wait();
      const uint32_t okm = c == 0 ? tmask[0] : (c == 1 ? tmask[1] : (c == 2 ? tmask[2] : tmask[3]));
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = ((okm >> i) & 1u) ? ex2f(fmaf(__uint_as_float(raw[i]), p.c, -mc)) : 0.f;
        const float p1 = ((okm >> (i + 1)) & 1u) ? ex2f(fmaf(__uint_as_float(raw[i + 1]), p.c, -mc)) : 0.f;
        l += p0 + p1;
        pk[i >> 1] = pack_bf16x2(p0, p1);
      }
      uint8_t* prow = sP + (c >> 1) * kTile + r * 128;  // key columns [64 * (c / 2), +64): one swizzle atom row
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = (c & 1) * 4 + i;
        *reinterpret_cast<uint4*>(prow + ((ch ^ (r & 7)) << 4)) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p_full);
    const float inv = 1.0f / l;
    const int t = r / p.G;
    const int s = blockIdx.x * p.G + rg;
    const bool valid = (t < p.T) && (s < p.S);
    bf16* op = p.out + (((long long)blockIdx.z * p.T + t) * p.S + s) * p.C + head * 64;
    mbar_wait(&o_full, 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld32(tO + lane_off + c * 32, raw);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          auto f = [&](int j) { return __uint_as_float(raw[16 * i + j]) * inv; };
          st_global_v8(op + c * 32 + 16 * i, pack_bf16x2(f(0), f(1)), pack_bf16x2(f(2), f(3)), pack_bf16x2(f(4), f(5)),
                       pack_bf16x2(f(6), f(7)), pack_bf16x2(f(8), f(9)), pack_bf16x2(f(10), f(11)),
                       pack_bf16x2(f(12), f(13)), pack_bf16x2(f(14), f(15)));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base_smem, 128);
  }
}

// ------------------------------------------------------------------------------------------
// Spatial attention, second generation (long sequences): one CTA owns TWO 128-query tiles of the
// same (frame, head) and keeps both in flight, so the tensor pipe works under the softmax of both
// (FlashAttention-4 style schedule, d = 64):
//   * warp 0 TMA (Q0, Q1 once; K / V blocks through a 4-slot ring in consumption order, each
//     loaded ONCE for both tiles), warp 1 MMA issuer, warps 2-5 softmax of tile 0, 6-9 of tile 1;
//   * TMEM (512 cols): S0 | S1 (128 fp32 cols each), O0 | O1 (64 each), P0 | P1 (64 each: 128 bf16
//     probabilities per row, two per column).  P never touches shared memory: the softmax warps
//     tcgen05.st it and P V is a TMEM-A tcgen05.mma — shared-memory bandwidth is left to K and V;
//   * a softmax thread owns one query row: it pulls the whole S row into registers and releases
//     S_t at once (s_free), so Q K^T of the next block runs under the exponentials of this one;
//   * O stays in TMEM for the whole key loop (the MMA accumulates across blocks); lazy rescale:
//     the running max only moves when it grew by more than 2^8, then the owning warp rescales its
//     O rows in place (tcgen05.ld / tcgen05.st) after P V of the previous block has completed.
// ------------------------------------------------------------------------------------------
constexpr int kAttn2Smem = (2 + 4) * kTile + 1024;  // Q0,Q1 | 4 K/V slots
constexpr int kAttn2Threads = 64 + 8 * 32;
#ifndef CTRLV_ATTN_EMU_EVERY
#define CTRLV_ATTN_EMU_EVERY 2
#endif
constexpr int kEmuEvery = CTRLV_ATTN_EMU_EVERY;  // every n-th pair of exponentials runs on the FMA pipe (0 = none);
// measured at S = 2560: none 380 us, 1/4 351, 1/3 351, 1/2 346, 2/3 372, 3/4 394, all 436

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__global__ void __launch_bounds__(kAttn2Threads, 1) attn2_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, s_full[2], s_free[2], p_full[2], o_full[2];
  __shared__ __align__(8) uint64_t kv_full[4], kv_empty[4];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                 // 2 tiles
  uint8_t* sKV = smem + 2 * kTile;    // 4 slots
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int nkv = p.nkv;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm);
    mbar_init(&q_full, 1);
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);
      mbar_init(&p_full[t], 4);
      mbar_init(&o_full[t], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  pdl_wait();
  const int q0 = blockIdx.x * 256;  // first query row of tile 0

  // K/V blocks travel through the ring in the order the MMA warp consumes them:
  //   K0, then for j = 0..nkv-1: [K(j+1)], V(j).
  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      const int colq = head * 64, colk = p.C + head * 64, colv = 2 * p.C + head * 64;
      mbar_expect_tx(&q_full, 2u * kTile);
      tma_load_3d(sQ, &p.tm, &q_full, colq, q0, blockIdx.z);
      tma_load_3d(sQ + kTile, &p.tm, &q_full, colq, q0 + 128, blockIdx.z);
      int n = 0;
      auto load = [&](int col, int blk) {
        const int slot = n & 3;
        mbar_wait_relaxed(&kv_empty[slot], (uint32_t)(((n >> 2) & 1) ^ 1));
        mbar_expect_tx(&kv_full[slot], (uint32_t)kTile);
        tma_load_3d(sKV + slot * kTile, &p.tm, &kv_full[slot], col, blk * 128, blockIdx.z);
        ++n;
      };
      load(colk, 0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) load(colk, j + 1);
        load(colv, j);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (warp-uniform, elected lane issues) ============
    const uint32_t idesc_qk = make_idesc(128, 128, 0, 0);
    const uint32_t idesc_pv = make_idesc(128, 64, 0, 1);  // A (= P) from TMEM, B (= V) MN-major
    const uint32_t aQ = smem_u32(sQ), aKV = smem_u32(sKV);
    // descriptor bases once (this warp's instruction stream is a single dependent chain: every descriptor rebuilt
    // inside a round is tensor-pipe idle time); a ring slot / k-step only adds to the >>4 address field
    const uint64_t dQ0 = make_sdesc(aQ, 16, 1024), dK0 = make_sdesc(aKV, 16, 1024), dV0 = make_sdesc(aKV, 1024, 1024);
    int n = 0;  // ring position of the block in use
    auto issue_qk = [&](int t, bool release) {  // S_t = Q_t K^T
      const int slot = n & 3;
      if (t == 0) {
        mbar_wait(&kv_full[slot], (uint32_t)((n >> 2) & 1));
        tc_fence_after();
      }
      if (elect_one()) {
        const uint64_t dq = dQ0 + (uint64_t)((t * kTile) >> 4), dk = dK0 + (uint64_t)((slot * kTile) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss(tmem_base + t * 128, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), idesc_qk, (uint32_t)(k != 0));
        if (release) umma_commit(&kv_empty[slot]);
        umma_commit(&s_full[t]);
      }
      __syncwarp();
    };
    auto issue_pv = [&](int t, int j, bool release) {  // O_t (+)= P_t V_j
      const int slot = n & 3;
      if (t == 0) {
        mbar_wait(&kv_full[slot], (uint32_t)((n >> 2) & 1));
        tc_fence_after();
      }
      if (elect_one()) {
        const uint64_t dv = dV0 + (uint64_t)((slot * kTile) >> 4);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_ts(tmem_base + 256 + t * 64, tmem_base + 384 + t * 64 + k * 8, dv + (uint64_t)(k * (2048 >> 4)), idesc_pv,
                  (uint32_t)((j | k) != 0));
        if (release) umma_commit(&kv_empty[slot]);
        umma_commit(&o_full[t]);
      }
      __syncwarp();
    };
    mbar_wait(&q_full, 0);
    issue_qk(0, false);
    issue_qk(1, true);
    ++n;
    for (int j = 0; j < nkv; ++j) {
      if (j + 1 < nkv) {
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&s_free[t], (uint32_t)(j & 1));  // S_t(j) is in the softmax warps' registers
          tc_fence_after();
          issue_qk(t, t == 1);
        }
        ++n;
      }
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], (uint32_t)(j & 1));  // P_t(j) in TMEM, O_t rescaled if needed
        tc_fence_after();
        issue_pv(t, j, t == 1);
      }
      ++n;
    }
  } else {
    // ============================ softmax warps =====================================
    const int t = (warp - 2) >> 2;                // tile
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                  // row within the tile
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 64 + lane_off;
    const uint32_t tP = tmem_base + 384 + t * 64 + lane_off;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&s_full[t], (uint32_t)(j & 1));
      tc_fence_after();
      uint32_t sreg[128];
      tmem_ld32(tS, sreg);
      tmem_ld32(tS + 32, sreg + 32);
      tmem_ld32(tS + 64, sreg + 64);
      tmem_ld32(tS + 96, sreg + 96);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[t]);
      const int kv0 = j * 128;
      const bool full_blk = (kv0 + 128 <= p.S);
      // ---- pass 1: row max
      float m_blk;
      if (full_blk) {
        float m0 = fmaxf(__uint_as_float(sreg[0]), __uint_as_float(sreg[1]));
        float m1 = fmaxf(__uint_as_float(sreg[2]), __uint_as_float(sreg[3]));
#pragma unroll
        for (int i = 4; i < 128; i += 4) {
          m0 = max3f(m0, __uint_as_float(sreg[i]), __uint_as_float(sreg[i + 1]));
          m1 = max3f(m1, __uint_as_float(sreg[i + 2]), __uint_as_float(sreg[i + 3]));
        }
        m_blk = fmaxf(m0, m1);
      } else {
        m_blk = -INFINITY;
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kv0 + i < p.S) m_blk = fmaxf(m_blk, __uint_as_float(sreg[i]));
      }
      // P_t and O_t are free once P V of the previous block has completed
      if (j > 0) {
        mbar_wait(&o_full[t], (uint32_t)((j - 1) & 1));
        tc_fence_after();
      }
      // ---- lazy rescale of the running max (and of O in TMEM) when it grew by more than 2^8
      if (j == 0) {
        m_run = m_blk;
      } else {
        const bool grow = (m_blk - m_run) * p.c > 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          const float m_new = grow ? m_blk : m_run;
          const float alpha = ex2f((m_run - m_new) * p.c);
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(tO + c * 16, o);
          }
          l_run *= alpha;
          m_run = m_new;
        }
      }
      const float mc = m_run * p.c;
      // ---- pass 2: P = exp2(s*c - m*c), packed bf16 pairs -> TMEM
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
          float p0, p1;
          if (kEmuEvery > 0 && ((i >> 1) % kEmuEvery) == kEmuEvery - 1) {
            // exp2 on the FMA pipe for a fraction of the pairs (the MUFU pipe is this kernel's floor):
            // x = n + f with n = round(x) taken from the mantissa of x + 1.5*2^23, 2^f by a cubic
            // (max rel. error 7.5e-5, below bf16), 2^n by adding n to the exponent field
            const uint64_t x2 = f2_fma(f2_pack(__uint_as_float(sreg[c * 64 + i]), __uint_as_float(sreg[c * 64 + i + 1])),
                                       f2_splat(p.c), f2_splat(-mc));
            float x0, x1;
            f2_unpack(x2, x0, x1);
            const uint64_t xc = f2_pack(fmaxf(x0, -126.0f), fmaxf(x1, -126.0f));
            const uint64_t tt = f2_add(xc, f2_splat(12582912.0f));
            const uint64_t ff = f2_fma(f2_add(tt, f2_splat(-12582912.0f)), f2_splat(-1.0f), xc);
            uint64_t q = f2_fma(ff, f2_splat(0.0551716676924127f), f2_splat(0.24261112208903293f));
            q = f2_fma(ff, q, f2_splat(0.6932609857127214f));
            q = f2_fma(ff, q, f2_splat(0.9999280735522258f));
            float q0, q1, t0, t1;
            f2_unpack(q, q0, q1);
            f2_unpack(tt, t0, t1);
            p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
            p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
          } else {
            p0 = ex2f(fmaf(__uint_as_float(sreg[c * 64 + i]), p.c, -mc));
            p1 = ex2f(fmaf(__uint_as_float(sreg[c * 64 + i + 1]), p.c, -mc));
          }
          if (!full_blk) {
            if (kv0 + c * 64 + i >= p.S) p0 = 0.f;
            if (kv0 + c * 64 + i + 1 >= p.S) p1 = 0.f;
          }
          l0 += p0; l1 += p1;
          pk[i >> 1] = pack_bf16x2(p0, p1);
        }
        tmem_st32(tP + c * 32, pk);
      }
      l_run += l0 + l1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    // ---- all P V of this tile done: normalise and store
    mbar_wait(&o_full[t], (uint32_t)((nkv - 1) & 1));
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int srow = q0 + t * 128 + r;
    const bool valid = srow < p.S;
    bf16* orow = p.out + ((long long)blockIdx.z * p.S + srow) * p.C + head * 64;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t raw[32];
      tmem_ld32(tO + c * 32, raw);
      tmem_ld_wait();
      if (valid) {
        bf16* op = orow + c * 32;  // whole 32-byte sectors per store (256-bit)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          auto f = [&](int j) { return __uint_as_float(raw[16 * i + j]) * inv; };
          st_global_v8(op + 16 * i, pack_bf16x2(f(0), f(1)), pack_bf16x2(f(2), f(3)), pack_bf16x2(f(4), f(5)),
                       pack_bf16x2(f(6), f(7)), pack_bf16x2(f(8), f(9)), pack_bf16x2(f(10), f(11)),
                       pack_bf16x2(f(12), f(13)), pack_bf16x2(f(14), f(15)));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool g_attn_init[64] = {false};  // per device ordinal
static int attn_init() {
  int dev = 0;
  CTRLV_CUDA(cudaGetDevice(&dev));
  CTRLV_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  if (g_attn_init[dev]) return CTRLV_OK;
  const int smem = 6 * kTile + 1024;
  CTRLV_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  CTRLV_CUDA(cudaFuncSetAttribute(tattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTattnSmem));
  CTRLV_CUDA(cudaFuncSetAttribute(attn2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttn2Smem));
  g_attn_init[dev] = true;
  return CTRLV_OK;
}

}  // namespace ctrlv

using namespace ctrlv;

extern "C" int ctrlv_attn_spatial(const void* qkv, int32_t frames, int32_t S, int32_t heads,
                                  float scale, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = attn_init();
  if (rc) return rc;
  CTRLV_CHECK_ARG(qkv && out && frames > 0 && S > 0 && heads > 0, "attn_spatial: bad arguments");
  CTRLV_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 31) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0,
                  "attn_spatial: out must be 32-byte aligned (256-bit stores), qkv 16-byte aligned");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.C = heads * 64; p.S = S; p.T = 1; p.G = 1; p.heads = heads;
  p.nkv = (S + 127) / 128;
  p.box_bytes = kTile;
  p.c = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(out);
  const uint64_t ld = 3ull * p.C;
  uint64_t dims[3] = {ld, (uint64_t)S, (uint64_t)frames};
  uint64_t strides[2] = {ld * 2, ld * 2 * (uint64_t)S};
  uint32_t box[3] = {64, 128, 1};
  rc = encode_tmap_bf16(&p.tm, qkv, 3, dims, strides, box, true);
  if (rc) return rc;
  // long sequences: two-tile kernel with P in TMEM; short ones: one tile per CTA, 2 CTAs per SM
  if (S >= 160) {
    dim3 grid2((S + 255) / 256, heads, frames);
    CTRLV_CUDA(launch_pdl(attn2_kernel, grid2, dim3(kAttn2Threads), (size_t)kAttn2Smem, stream, p));
    return CTRLV_OK;
  }
  dim3 grid((S + 127) / 128, heads, frames);
  CTRLV_CUDA(launch_pdl(attn_kernel, grid, dim3(kAttnThreads), (size_t)(6 * kTile + 1024), stream, p));
  return CTRLV_OK;
}

extern "C" int ctrlv_attn_temporal(const void* qkv, int32_t B, int32_t T, int32_t S, int32_t heads,
                                   float scale, void* out, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  int rc = attn_init();
  if (rc) return rc;
  CTRLV_CHECK_ARG(qkv && out && B > 0 && S > 0 && heads > 0, "attn_temporal: bad arguments");
  CTRLV_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 31) == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0,
                  "attn_temporal: out must be 32-byte aligned (256-bit stores), qkv 16-byte aligned");
  CTRLV_CHECK_ARG(T >= 1 && T <= 128, "attn_temporal: T=%d unsupported (1..128)", T);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.C = heads * 64; p.S = S; p.T = T; p.heads = heads;
  p.G = 128 / T;
  if (p.G > S) p.G = S;
  p.nkv = 1;
  p.box_bytes = 64 * p.G * T * 2;
  p.c = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<bf16*>(out);
  const uint64_t ld = 3ull * p.C;
  uint64_t dims[4] = {ld, (uint64_t)S, (uint64_t)T, (uint64_t)B};
  uint64_t strides[3] = {ld * 2, ld * 2 * (uint64_t)S, ld * 2 * (uint64_t)S * (uint64_t)T};
  uint32_t box[4] = {64, (uint32_t)p.G, (uint32_t)T, 1};
  rc = encode_tmap_bf16(&p.tm, qkv, 4, dims, strides, box, true);
  if (rc) return rc;
  dim3 grid((S + p.G - 1) / p.G, heads, B);
  CTRLV_CUDA(launch_pdl(tattn_kernel, grid, dim3(kTattnThreads), (size_t)kTattnSmem, stream, p));
  return CTRLV_OK;
}
