// tcgen05 implicit-GEMM contraction: the one tensor-core kernel behind every Linear, 3x3 conv,
// stride-2 conv, skip-concat conv, fused 1x1 shortcut and (3,1,1) temporal conv of the
// Box2Video denoise step.
//
//   D[m][n] = sum over K segments s, channels c:  A_s[coord(m) + shift_s][c] * W[n][k(s, c)]
//
// * A operand: channels-last activations, fetched by TMA as 4-D boxes (64 channels x bx x by x bz
//   sites = up to 128 rows) straight into 128B-swizzled shared memory; conv halos, temporal halos
//   and ragged tiles are TMA out-of-bounds zero fill — no im2col, no padding copies.
// * B operand: weights [N][K] (K contiguous), 2-D TMA boxes of 64 x BN.
// * tcgen05.mma (cta_group::1, M=128, N=BN, K=16) issued by one thread, fp32 accumulators in
//   TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the main loop of
//   tile i+1.  Persistent CTAs (one per SM), static round-robin tile schedule.
// * Warp roles (16 warps = 4 warpgroups): warpgroup 0 = {TMA producer, MMA issuer (+ TMEM alloc), two
//   idle warps}, warpgroups 1-3 = 12 epilogue warps (each owns the TMEM lane quarter warp_id % 4, one
//   accumulator row per thread; the three warps of a quarter split the 32-column chunks).  After the
//   set-up barrier warpgroup 0 shrinks to 40 registers per thread and the epilogue warpgroups grow to
//   152 (setmaxnreg): the epilogue holds an accumulator chunk, a prefetched residual chunk and its
//   row bookkeeping in registers without spilling.
// * Optional stream-K schedule (igemm_kernel<CG, true> + igemm_fixup_kernel, ctrlv_epilogue.splitk_ws): contiguous
//   k-block ranges per CTA instead of whole tiles, fp32 partial tiles in the caller's workspace, a fix-up launch that
//   adds them in contributor order and runs the same fused epilogue.  Off on the denoise step's shapes (measured
//   level, see sk_wanted()).
#include <cstdlib>

#include "common.cuh"
#include "fastdiv.h"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

#ifndef CTRLV_EPI_ROWOWN
#define CTRLV_EPI_ROWOWN 1
#endif
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kMaxStages = 8;
#ifndef CTRLV_MMA_BATCH
#define CTRLV_MMA_BATCH 4
#endif
constexpr int kMmaBatch = CTRLV_MMA_BATCH;  // most k-blocks issued per round of the MMA warp (IgemmParams::mma_batch)
constexpr int kEpiWarps = 12;     // epilogue warps: 3 per TMEM lane quarter
constexpr int kFirstEpiWarp = 4;  // warps 0-3: TMA producer, MMA issuer, two idle (one warpgroup)
constexpr int kThreads = (kFirstEpiWarp + kEpiWarps) * 32;

struct IgemmSeg {
  int map, c0, nchunk, dx, dy, dz;
};

struct IgemmParams {
  CUtensorMap tmA[CTRLV_MAX_SRC];
  CUtensorMap tmB;
  IgemmSeg seg[CTRLV_MAX_SEG];
  int nseg, kblocks;
  int X, Y, Z, bx, by, bz;
  int tiles_x, tiles_y, tiles_z, tiles_n, tiles_total;
  FastDiv fd_n, fd_x, fd_y;  // division by tiles_n / tiles_x / tiles_y
  int BN, N, stages, a_bytes, stage_bytes, tmem_cols;
  int omx, omy, oox, ooy, oX, oY;  // output-row remap (identity: 1, 1, 0, 0, X, Y)
  int cg;  // 1 or 2 CTAs per tile (tcgen05 cta_group)
  int m_tiles;    // tiles_x * tiles_y * tiles_z
  int mma_batch;  // k-blocks per round of the MMA warp (<= kMmaBatch, <= stages)
  FastDiv fd_rpu, fd_cg;  // GroupNorm statistics: division by rows per unit / channels per group
  // stream-K schedule (igemm_kernel<CG, true>): the tiles_total * kblocks k-blocks of the problem, tile-major, are
  // dealt to the CTAs (pairs) as contiguous ranges of sk_per k-blocks
  int sk_per, sk_total;
  FastDiv fd_kb, fd_per;  // division by kblocks / sk_per
  float* sk_part;         // fp32 partial tiles [gridDim.x * 2][128][BN]
  ctrlv_epilogue ep;
};

// ---- work schedule ------------------------------------------------------------------------------
// SK = false: whole tiles, dealt round-robin to the CTAs (pairs).  SK = true (stream-K): CTA g owns the k-blocks
// [g * sk_per, (g + 1) * sk_per) of the tile-major k-block sequence, i.e. the tail of one tile, whole tiles, the
// head of another — every SM gets the same amount of tensor work whatever the tile count.
template <bool SK>
struct WorkIter {
  int a, b, c;  // SK: position, end of the range, -;  !SK: next tile, stride, tiles_total
  __device__ __forceinline__ void init(const IgemmParams& p, int g, int ng) {
    if (SK) {
      a = g * p.sk_per;
      b = min(p.sk_total, a + p.sk_per);
      c = 0;
    } else {
      a = g; b = ng; c = p.tiles_total;
    }
  }
  // next unit of work: k-blocks [kb0, kb1) of `tile`
  __device__ __forceinline__ bool next(const IgemmParams& p, int& tile, int& kb0, int& kb1) {
    if (SK) {
      if (a >= b) return false;
      tile = (int)fd_div((uint32_t)a, p.fd_kb);
      kb0 = a - tile * p.kblocks;
      kb1 = min(p.kblocks, kb0 + (b - a));
      a += kb1 - kb0;
      return true;
    } else {
      if (a >= c) return false;
      tile = a; kb0 = 0; kb1 = p.kblocks;
      a += b;
      return true;
    }
  }
};

__device__ __forceinline__ int rowbias_index(const ctrlv_epilogue& ep, int m) {
  int a = m / ep.rb_div;
  if (ep.rb_mode == 1) return a;
  if (ep.rb_mode == 2) return a % ep.rb_mod;
  return (a * ep.rb_mod + m % ep.rb_mod + ep.rb_off) % ep.rb_B;
}

// residual rows of one chunk, fetched BEFORE the TMEM load so their latency overlaps it
// ---- epilogue data movement -------------------------------------------------------------------
// A thread owns one accumulator ROW, so its natural global accesses are 32 scattered 16-byte pieces
// per warp instruction — the LSU, not HBM, becomes the limit for short-K GEMMs.  Every bf16
// residual read / output write is therefore transposed through a 2 KB warp-private smem tile so that
// a warp instruction touches whole 64-byte row segments (CPR lanes per row): 4x fewer LSU
// transactions, no cross-warp synchronisation.
template <int NV>
struct EpRows {  // per (tile, warp): the rows this lane touches in the transposed (coalesced) mapping
  static constexpr int CPR = NV / 8;    // 16-byte pieces per row of one chunk
  static constexpr int RPI = 32 / CPR;  // rows covered by one warp-wide 16-byte access
  int mT[CPR];  // output row (>= 0), or -1 for a row outside the problem (one register instead of two)
  __device__ __forceinline__ void init(long long m, bool valid) {
    const int lane = threadIdx.x & 31;
    const int mv = valid ? (int)m : -1;
#pragma unroll
    for (int i = 0; i < CPR; ++i) mT[i] = __shfl_sync(0xffffffffu, mv, lane / CPR + RPI * i);
  }
  // 16-byte slot of (row, piece) in the warp tile, XOR-swizzled so that both the row-owner access
  // (lane = row) and the transposed access (CPR lanes per row) are bank-conflict free
  static __device__ __forceinline__ int slot(int row, int piece) {
    return row * CPR + (CPR == 4 ? (piece ^ ((row >> 1) & 3)) : (piece ^ ((row >> 2) & 1)));
  }
};

// ---- GroupNorm statistics of the stored tile (ep.gn_sums) ---------------------------------------
// The nn.GroupNorm that consumes this kernel's output needs (sum, sum of squares) per (statistics unit,
// channel group).  The epilogue already holds every output value: after the transposed read of the
// staged bf16 tile a lane owns 8 consecutive columns of 4 rows, so it sums those rows per column,
// splits the 8 columns between the (at most two) groups they fall in, an xor-shuffle reduces over
// the 8 lanes that hold the same columns, and one lane per column piece adds the partials to
// gn_sums[unit][group] as 64-bit FIXED-POINT integers (scale 2^16): integer addition is associative, so
// the result is bit-identical from run to run whatever order the atomics land in — no float atomics.
constexpr float kGnFix = 65536.0f;
struct GnTile {  // per (tile, warp)
  int uT[4];     // statistics unit of the 4 rows this lane touches in the transposed mapping (-1: no row)
  int u0;        // smallest unit among the warp's rows (-1: the warp has no valid row)
  bool multi;    // the warp's 32 rows straddle more than one unit (small frames): one pass per unit
  __device__ __forceinline__ void init(const IgemmParams& p, long long m, bool valid) {
    const int lane = threadIdx.x & 31;
    const int mu = valid ? (int)fd_div((uint32_t)m, p.fd_rpu) : -1;
#pragma unroll
    for (int i = 0; i < 4; ++i) uT[i] = __shfl_sync(0xffffffffu, mu, lane / 4 + 8 * i);
    int mn = mu < 0 ? 0x7fffffff : mu, mx = mu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    u0 = mx < 0 ? -1 : mn;
    multi = mx != u0;
  }
};

__device__ __forceinline__ void gn_atomic(void* sums, size_t rep_off, int unit, int group, float s, float q) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(sums) + rep_off + ((size_t)unit * 32 + group) * 2;
  atomicAdd(a, (unsigned long long)__float2ll_rn(s * kGnFix));
  atomicAdd(a + 1, (unsigned long long)__float2ll_rn(q * kGnFix));
}

// ut[i]: the 8 stored bf16 values (columns o0 + 8*(lane%4) ..) of row lane/4 + 8*i
__device__ __forceinline__ void gn_accumulate(const uint4* ut, const GnTile& gn, const ctrlv_epilogue& ep,
                                              const IgemmParams& p, int o0) {
  const int lane = threadIdx.x & 31;
  const int c0 = ep.gn_c_off + o0 + (lane & 3) * 8;  // channel of this lane's first column in the consumer's numbering
  const int gA = (int)fd_div((uint32_t)c0, p.fd_cg);
  const int jb = (gA + 1) * ep.gn_cg - c0;           // columns j >= jb fall into group gA + 1
  // The table is replicated gn_rep times (a power of two) and every warp adds into "its" replica: with only
  // a few statistics units (a temporal norm has one per clip) all SMs would otherwise hammer the same
  // handful of L2 lines (measured: +33 us on a 62 us launch).  The consumer sums the replicas.
  const size_t rep_off = (size_t)((blockIdx.x * (unsigned)kEpiWarps + (threadIdx.x >> 5)) & (unsigned)(ep.gn_rep - 1)) *
                         (size_t)ep.gn_units * 64;
  int ucur = gn.u0;
  while (true) {
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] = 0.f; q[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (gn.uT[i] == ucur) {
        const uint32_t w[4] = {ut[i].x, ut[i].y, ut[i].z, ut[i].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(w[k]);
          s[2 * k] += f.x; q[2 * k] = fmaf(f.x, f.x, q[2 * k]);
          s[2 * k + 1] += f.y; q[2 * k + 1] = fmaf(f.y, f.y, q[2 * k + 1]);
        }
      }
    }
    float sA = 0.f, qA = 0.f, sB = 0.f, qB = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < jb) { sA += s[j]; qA += q[j]; } else { sB += s[j]; qB += q[j]; }
    }
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      sA += __shfl_xor_sync(0xffffffffu, sA, o); qA += __shfl_xor_sync(0xffffffffu, qA, o);
      sB += __shfl_xor_sync(0xffffffffu, sB, o); qB += __shfl_xor_sync(0xffffffffu, qB, o);
    }
    if (lane < 4) {
      gn_atomic(ep.gn_sums, rep_off, ucur, gA, sA, qA);
      if (jb < 8) gn_atomic(ep.gn_sums, rep_off, ucur, gA + 1, sB, qB);
    }
    if (!gn.multi) break;
    int nx = 0x7fffffff;  // next statistics unit present in this warp's rows
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (gn.uT[i] > ucur) nx = min(nx, gn.uT[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nx = min(nx, __shfl_xor_sync(0xffffffffu, nx, o));
    if (nx == 0x7fffffff) break;
    ucur = nx;
  }
}

// Prefetch loads are volatile asm: the compiler may not sink them below the (volatile) accumulator wait,
// so their latency overlaps the MMA instead of sitting on the epilogue's critical path.
__device__ __forceinline__ float ldg_f32_early(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_v4_early(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ void ldg_v8_early(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}

template <int NV>
struct ResPrefetch {
  static constexpr int CPR = NV / 8;
  uint4 r1[CPR];  // transposed mapping: piece (lane % CPR) of rows lane / CPR + RPI * i
                  // (row-owner mapping: the CPR pieces of this lane's own row)
  bool full;      // warp-uniform: the whole chunk lies inside the stored columns
  __device__ __forceinline__ void issue(const ctrlv_epilogue& ep, const EpRows<NV>& rows, int o0, int n_store,
                                        int n0, int N, bool rowown, long long m, bool valid) {
    const int lane = threadIdx.x & 31;
    full = (o0 + NV <= n_store) && (n0 < N);
    if (NV == 32 && rowown) {
      if (full && ep.res1 && valid) {
        const bf16* rp = reinterpret_cast<const bf16*>(ep.res1) + (size_t)m * ep.ld_res1 + o0;
        ldg_v8_early(rp, r1[0], r1[1]);
        ldg_v8_early(rp + 16, r1[CPR > 2 ? 2 : 0], r1[CPR > 3 ? 3 : 1]);
      }
      return;
    }
    if (full && ep.res1) {
#pragma unroll
      for (int i = 0; i < CPR; ++i) {
        r1[i] = make_uint4(0, 0, 0, 0);
        if (rows.mT[i] >= 0)
          r1[i] = ldg_v4_early(reinterpret_cast<const bf16*>(ep.res1) + (size_t)rows.mT[i] * ep.ld_res1 + o0 +
                               (lane % CPR) * 8);
      }
    }
  }
};

// bias (+ the warp-uniform rowbias row) of the up-to-three chunks a warp serves: lane j holds column
// n_base + c*32 + j; fetched before the accumulator wait, redistributed through smem per chunk
struct BiasPrefetch {
  float b[3], u[3];
  __device__ __forceinline__ void issue(const ctrlv_epilogue& ep, int n_base, int N, int nch, int sub,
                                        const float* rb_uniform) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int c = sub + 3 * i;
      const int nl = n_base + c * 32 + lane;
      b[i] = 0.f; u[i] = 0.f;
      if (c < nch && nl < N) {
        if (ep.bias) b[i] = ldg_f32_early(ep.bias + nl);
        if (rb_uniform) u[i] = ldg_f32_early(rb_uniform + nl);
      }
    }
  }
};

__device__ __forceinline__ void add_bf16x8(float* v, const uint4& u, float s) {  // v += s * u, packed fp32x2
  float2 f;
  f = unpack_bf16x2(u.x); v[0] += s * f.x; v[1] += s * f.y;
  f = unpack_bf16x2(u.y); v[2] += s * f.x; v[3] += s * f.y;
  f = unpack_bf16x2(u.z); v[4] += s * f.x; v[5] += s * f.y;
  f = unpack_bf16x2(u.w); v[6] += s * f.x; v[7] += s * f.y;
}

// scale, add residual streams, convert and store NV consecutive outputs of row m starting at output
// column o0.  Executed by ALL lanes of the warp (the smem transposes are warp-collective).
template <int NV>
__device__ __forceinline__ void ep_finish(float* v, const ctrlv_epilogue& ep, long long m, bool valid, int o0,
                                          int n_store, const ResPrefetch<NV>& pf, const EpRows<NV>& rows,
                                          uint4* wst, const GnTile& gn, const IgemmParams& p, bool rowown) {
  constexpr int CPR = NV / 8;
  constexpr int RPI = 32 / CPR;
  const int lane = threadIdx.x & 31;
  if (ep.s_acc != 1.0f) {  // warp-uniform
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] *= ep.s_acc;
  }
  if (pf.full) {
    if (ep.res1 && NV == 32 && rowown) {
      // row-owner mapping: the lane's own 64 bytes arrived as two whole-sector 256-bit loads
      if (valid) {
#pragma unroll
        for (int j = 0; j < CPR; ++j) add_bf16x8(v + 8 * j, pf.r1[j], ep.s_res1);
      }
    } else if (ep.res1) {
      __syncwarp();
#pragma unroll
      for (int i = 0; i < CPR; ++i) wst[EpRows<NV>::slot(lane / CPR + RPI * i, lane % CPR)] = pf.r1[i];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < CPR; ++j) add_bf16x8(v + 8 * j, wst[EpRows<NV>::slot(lane, j)], ep.s_res1);
    }
    if (ep.res2 && valid) {
      const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(ep.res2) + (size_t)m * ep.ld_res2 + o0);
      if ((reinterpret_cast<uintptr_t>(rp) & 31) == 0) {  // whole 32-byte sectors per load
#pragma unroll
        for (int j = 0; j < CPR; j += 2) {
          uint4 lo, hi;
          ld_global_nc_v8(rp + j, lo, hi);
          add_bf16x8(v + 8 * j, lo, ep.s_res2);
          add_bf16x8(v + 8 * j + 8, hi, ep.s_res2);
        }
      } else {
#pragma unroll
        for (int j = 0; j < CPR; ++j) add_bf16x8(v + 8 * j, __ldg(rp + j), ep.s_res2);
      }
    }
    if (ep.out_f32 && valid) {  // (before the bf16 path: v[] is dead once it is packed)
      float4* op = reinterpret_cast<float4*>(ep.out_f32 + (size_t)m * ep.ld_out_f32 + o0);
#pragma unroll
      for (int j = 0; j < NV; j += 4) op[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (ep.out && NV == 32 && rowown) {
      // no GroupNorm statistics to take from the staged tile: store the lane's own row as two whole 32-byte
      // sectors (256-bit stores) — no smem transposes, no per-row address arithmetic for four rows
      if (valid) {
        bf16* op = reinterpret_cast<bf16*>(ep.out) + (size_t)m * ep.ld_out + o0;
#pragma unroll
        for (int h = 0; h < NV / 16; ++h)
          st_global_v8(op + 16 * h, pack_bf16x2(v[16 * h], v[16 * h + 1]), pack_bf16x2(v[16 * h + 2], v[16 * h + 3]),
                       pack_bf16x2(v[16 * h + 4], v[16 * h + 5]), pack_bf16x2(v[16 * h + 6], v[16 * h + 7]),
                       pack_bf16x2(v[16 * h + 8], v[16 * h + 9]), pack_bf16x2(v[16 * h + 10], v[16 * h + 11]),
                       pack_bf16x2(v[16 * h + 12], v[16 * h + 13]), pack_bf16x2(v[16 * h + 14], v[16 * h + 15]));
      }
    } else if (ep.out && NV == 16) {
      // GEGLU chunks yield only 32 B per row: the transpose does not pay (measured), store directly
      // One 256-bit store per lane = one whole 32-byte sector (two 16-byte stores leave the L2 with
      // half-written sectors: measured, the store path cost a quarter of the K = 320 GEGLU kernel).
      if (valid) {
        bf16* op = reinterpret_cast<bf16*>(ep.out) + (size_t)m * ep.ld_out + o0;
        if ((reinterpret_cast<uintptr_t>(op) & 31) == 0) {
          st_global_v8(op, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                       pack_bf16x2(v[6], v[7]), pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                       pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
        } else {
          uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
          for (int j = 0; j < CPR; ++j)
            o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
        }
      }
    } else if (ep.out) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < CPR; ++j) {
        uint4 u;
        u.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
        u.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
        u.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
        wst[EpRows<NV>::slot(lane, j)] = u;
      }
      __syncwarp();
      uint4 ut[CPR];
#pragma unroll
      for (int i = 0; i < CPR; ++i) {
        ut[i] = wst[EpRows<NV>::slot(lane / CPR + RPI * i, lane % CPR)];
        if (rows.mT[i] >= 0)
          *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(ep.out) + (size_t)rows.mT[i] * ep.ld_out + o0 +
                                    (lane % CPR) * 8) = ut[i];
      }
      if (NV == 32 && ep.gn_sums && gn.u0 >= 0) gn_accumulate(ut, gn, ep, p, o0);  // warp-uniform
    }
  } else if (valid) {
    // ragged tail of a padded-N problem (e.g. conv_out with 4 real channels): scalar, predicated
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int o = o0 + j;
      if (o < n_store) {
        float t = v[j];
        if (ep.res1)
          t += ep.s_res1 * __bfloat162float(reinterpret_cast<const bf16*>(ep.res1)[(size_t)m * ep.ld_res1 + o]);
        if (ep.res2)
          t += ep.s_res2 * __bfloat162float(reinterpret_cast<const bf16*>(ep.res2)[(size_t)m * ep.ld_res2 + o]);
        if (ep.out) reinterpret_cast<bf16*>(ep.out)[(size_t)m * ep.ld_out + o] = __float2bfloat16(t);
        if (ep.out_f32) ep.out_f32[(size_t)m * ep.ld_out_f32 + o] = t;
      }
    }
  }
}

constexpr int kSkMaxContrib = 4;  // most CTAs (pairs) that share one tile (the launcher sizes the k-ranges accordingly)
// 256-bit load that observes other SMs' stores of this launch (L2, never L1); p 32-byte aligned
__device__ __forceinline__ void ld_global_coherent_v8(const void* p, uint32_t* v) {
  asm volatile("ld.relaxed.gpu.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "l"(p)
               : "memory");
}

// stream-K: where the fp32 partial tiles of one output tile lie (one slot per contributing CTA, see igemm_kernel)
struct SkParts {
  const float* base;  // p.sk_part + this thread's (chunk-independent) row offset
  size_t slot_elems;  // 128 * BN
  int gf, gl;         // first / last contributing CTA (pair)
  int w_first;        // slot parity of contributor gf (every later contributor starts inside the tile: parity 0)
  int cg, crank;
  __device__ __forceinline__ const float* slot(int g) const {
    return base + (size_t)((g * 2 + (g == gf ? w_first : 0)) * cg + crank) * slot_elems;
  }
};

// one 32-column accumulator chunk of one row: TMEM load, bias, (GEGLU), residual, store
// WS: the accumulator is the sum of the tile's stream-K partials (fixed contributor order: reproducible) instead
template <bool GEGLU, bool WS = false>
__device__ __forceinline__ void ep_chunk(const ctrlv_epilogue& ep, uint32_t taddr, long long m, bool valid, int n0,
                                         int n_store, float* sb, uint4* wst, const float* rb, float bv,
                                         const ResPrefetch<GEGLU ? 16 : 32>& pf,
                                         const EpRows<GEGLU ? 16 : 32>& rows, const GnTile& gn, const IgemmParams& p,
                                         bool rowown, const SkParts* sk = nullptr, int sk_off = 0) {
  constexpr int NV = GEGLU ? 16 : 32;
  uint32_t raw[32];
  if (WS) {
    // two half-chunks of 16 columns; the loads of ALL contributors of a half (<= kSkMaxContrib x two 256-bit
    // loads, whole sectors, L2-coherent) are in flight together — the fix-up is a chain of L2 round trips
    const int nc = sk->gl - sk->gf + 1;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint32_t t[kSkMaxContrib][16];
#pragma unroll
      for (int i = 0; i < kSkMaxContrib; ++i) {
        if (i < nc) {
          const float* src = sk->slot(sk->gf + i) + sk_off + 16 * h;
          ld_global_coherent_v8(src, t[i]);
          ld_global_coherent_v8(src + 8, t[i] + 8);
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a = __uint_as_float(t[0][j]);  // contributor order: the sum is the same from run to run
#pragma unroll
        for (int i = 1; i < kSkMaxContrib; ++i)
          if (i < nc) a += __uint_as_float(t[i][j]);
        raw[16 * h + j] = __float_as_uint(a);
      }
    }
  } else {
    tmem_ld32(taddr, raw);
  }
  // bias (+ warp-uniform rowbias): lane j fetched column j; broadcast through this warp's smem row
  __syncwarp();  // previous chunk's readers are done with the row
  sb[threadIdx.x & 31] = bv;
  __syncwarp();
  if (!WS) tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
#pragma unroll
  for (int j = 0; j < 32; j += 4) {  // packed fp32x2 adds: 16 instead of 32 issue slots
    const float4 b = *reinterpret_cast<const float4*>(sb + j);
    v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
  }
  if (rb) {  // (nullptr for rows outside the problem)
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(rb + n0 + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (GEGLU) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      float o0, o1;
      geglu2_f(v[2 * j], v[2 * j + 1], v[2 * j + 2], v[2 * j + 3], o0, o1);
      v[j] = o0; v[j + 1] = o1;
    }
  }
  ep_finish<NV>(v, ep, m, valid, GEGLU ? (n0 >> 1) : n0, n_store, pf, rows, wst, gn, p, rowown);
}

// All 32-column chunks of one accumulator row owned by this warp (c = sub, sub + G, sub + 2G, G = 3
// chunk groups, at most 3 chunks for BN <= 256).  Residual rows are fetched one chunk AHEAD — the
// first one before the accumulator is even ready — so their HBM latency overlaps the MMA wait and
// the previous chunk's math instead of sitting on the critical path.
template <bool GEGLU>
__device__ __forceinline__ void ep_tile(const ctrlv_epilogue& ep, uint64_t* tfull, uint32_t tphase, uint32_t t_row,
                                        long long m, int n_base, int N, int BN, int n_store, bool valid, int sub,
                                        float* sbias, uint4* wst, const float* rb, const float* rb_uniform,
                                        const IgemmParams& p, bool rowown) {
  constexpr int NV = GEGLU ? 16 : 32;
  constexpr int G = kEpiWarps / 4;
  const int nch = BN / 32;
  const int c0 = sub, c1 = sub + G, c2 = sub + 2 * G;
  EpRows<NV> rows;
  rows.init(m, valid);
  GnTile gn;
  gn.u0 = -1; gn.multi = false;
  if (!GEGLU && ep.gn_sums) gn.init(p, m, valid);
  ResPrefetch<NV> pa, pb;
  BiasPrefetch bp;
  auto o_of = [&](int c) { return GEGLU ? ((n_base + c * 32) >> 1) : (n_base + c * 32); };
  // every bias value and the first TWO chunks' residual rows are requested before the accumulator is
  // ready: their latency hides under the MMA wait instead of under epilogue math
  bp.issue(ep, n_base, N, nch, sub, rb_uniform);
  if (c0 < nch) pa.issue(ep, rows, o_of(c0), n_store, n_base + c0 * 32, N, rowown, m, valid);
  if (c1 < nch) pb.issue(ep, rows, o_of(c1), n_store, n_base + c1 * 32, N, rowown, m, valid);
  mbar_wait(tfull, tphase);
  tc_fence_after();
  if (c0 >= nch) return;
  __syncwarp();
  ep_chunk<GEGLU>(ep, t_row + (uint32_t)(c0 * 32), m, valid, n_base + c0 * 32, n_store, sbias, wst, rb,
                  bp.b[0] + bp.u[0], pa, rows, gn, p, rowown);
  if (c1 >= nch) return;
  if (c2 < nch) pa.issue(ep, rows, o_of(c2), n_store, n_base + c2 * 32, N, rowown, m, valid);
  __syncwarp();
  ep_chunk<GEGLU>(ep, t_row + (uint32_t)(c1 * 32), m, valid, n_base + c1 * 32, n_store, sbias, wst, rb,
                  bp.b[1] + bp.u[1], pb, rows, gn, p, rowown);
  if (c2 >= nch) return;
  __syncwarp();
  ep_chunk<GEGLU>(ep, t_row + (uint32_t)(c2 * 32), m, valid, n_base + c2 * 32, n_store, sbias, wst, rb,
                  bp.b[2] + bp.u[2], pa, rows, gn, p, rowown);
}

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) shares
// one 256-row x BN tile: each CTA loads its own 128 A rows and HALF of the B tile, the leader issues
// M=256 MMAs that read both halves — halves the shared-memory and L2 traffic of the B operand.
template <int CG, bool SK>
__global__ void __launch_bounds__(kThreads, 1) igemm_kernel(const __grid_constant__ IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tfull_bar[2];
  __shared__ __align__(8) uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float wbias[kEpiWarps][32];   // per-warp broadcast row for the bias chunk
  __shared__ __align__(16) uint4 wstage[kEpiWarps][128];  // per-warp 32 x 64 B transpose tile

  // 1024-byte aligned tile ring
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();  // the next kernel may be scheduled as SMs drain; it blocks in its own pdl_wait()

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nseg; ++i) tma_prefetch_desc(&p.tmA[p.seg[i].map]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], kEpiWarps * CG);  // CG = 2: both CTAs' epilogue warps free the leader's buffer
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc(&tmem_base_smem, (uint32_t)p.tmem_cols);
    else tmem_alloc_cg2(&tmem_base_smem, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  pdl_wait();  // everything above overlapped the predecessor's tail; its results are needed from here

  // work unit: (super-tile of CG consecutive m-tiles, n-tile); units are dealt round-robin to
  // clusters; inside a pair CTA r owns m-tile CG*super + r
  const int unit0 = (int)blockIdx.x / CG;
  const int nunits = (int)gridDim.x / CG;
  // register reallocation between the warpgroups: each warpgroup executes ONE setmaxnreg at the head of
  // its role branch (ptxas budgets the code below it accordingly)
  if (warp < kFirstEpiWarp) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ================================ TMA producer ================================
    {  // warp-uniform loop, one elected lane issues (see the MMA issuer below)
      int stage = 0;
      uint32_t phase = 0;
      WorkIter<SK> wi;
      wi.init(p, unit0, nunits);
      int tile, kb0, kb1;
      while (wi.next(p, tile, kb0, kb1)) {
        uint32_t q_, nt_, tx_, ty_, tz_;
        fd_divmod((uint32_t)tile, p.fd_n, q_, nt_);
        fd_divmod(q_ * CG + crank, p.fd_x, q_, tx_);
        fd_divmod(q_, p.fd_y, tz_, ty_);
        const int nt = (int)nt_, tx = (int)tx_, ty = (int)ty_;
        const int tz = (int)tz_;  // may run past tiles_z for the odd tail: TMA zero-fills
        const int x0 = tx * p.bx, y0 = ty * p.by, z0 = tz * p.bz;
        int kbs = 0;  // first k-block of segment s
        for (int s = 0; s < p.nseg; ++s) {
          const IgemmSeg sg = p.seg[s];
          const int ch0 = SK ? max(0, kb0 - kbs) : 0;
          const int ch1 = SK ? min(sg.nchunk, kb1 - kbs) : sg.nchunk;
          for (int ch = ch0; ch < ch1; ++ch) {
            const int kb = kbs + ch;
            mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sb = sa + kBM * kBK * 2;
            if (elect_one()) {
              if (CG == 2) {
                // both CTAs credit the LEADER's full barrier; the leader expects the pair's bytes
                const uint32_t lead_bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
                if (crank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(2 * (p.a_bytes + (p.BN / 2) * kBK * 2)));
                tma_load_4d_cg2(sa, &p.tmA[sg.map], lead_bar, sg.c0 + ch * kBK, x0 + sg.dx, y0 + sg.dy, z0 + sg.dz);
                tma_load_2d_cg2(sb, &p.tmB, lead_bar, kb * kBK, nt * p.BN + (int)crank * (p.BN / 2));
              } else {
                mbar_expect_tx(&full_bar[stage], (uint32_t)(p.a_bytes + p.BN * kBK * 2));
                tma_load_4d(sa, &p.tmA[sg.map], &full_bar[stage], sg.c0 + ch * kBK, x0 + sg.dx, y0 + sg.dy, z0 + sg.dz);
                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * kBK, nt * p.BN);
              }
            }
            __syncwarp();
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
          }
          kbs += sg.nchunk;
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (crank == 0) {
      // Warp-uniform loop: all lanes wait on the barriers, ONE elected lane issues.  (Issuing from a
      // `lane == 0` branch makes ptxas wrap every UTCHMMA / UTCBAR in an ELECT..BRA.U.ANY waterfall
      // loop, which costs more issue time than the MMAs themselves for narrow tiles.)
      const uint32_t idesc = make_idesc(kBM * CG, p.BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      WorkIter<SK> wi;
      wi.init(p, unit0, nunits);
      int tile, kb0, kb1;
      for (int it = 0; wi.next(p, tile, kb0, kb1); ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait_relaxed(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.BN);
        // kMmaBatch k-blocks per round of this warp: one barrier-wait / elect / branch sequence (~75 instructions,
        // 400-600 cycles of a single warp's dependent uniform-datapath work) then carries 4 * kMmaBatch MMAs.
        // With one k-block per round the issue stream outlasted the tensor work of every tile narrower than 256
        // columns (a 64-wide k-block of a 128-column tile is 256 tensor cycles).
        for (int kb = kb0; kb < kb1; kb += p.mma_batch) {
          const int nb = min(p.mma_batch, kb1 - kb);
          int st[kMmaBatch];
          {
            int s_ = stage;
            uint32_t ph_ = phase;
#pragma unroll
            for (int b = 0; b < kMmaBatch; ++b) {
              st[b] = s_;
              if (b < nb) mbar_wait(&full_bar[s_], ph_);
              if (++s_ == p.stages) { s_ = 0; ph_ ^= 1; }
            }
          }
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int b = 0; b < kMmaBatch; ++b) {
              if (b < nb) {
                const uint32_t sa = smem_u32(smem + (size_t)st[b] * p.stage_bytes);
                const uint32_t sb = sa + kBM * kBK * 2;
                const uint64_t da = make_sdesc(sa, 16, 1024);
                const uint64_t db = make_sdesc(sb, 16, 1024);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                  // advance 32 bytes (16 bf16) inside the 128B swizzle atom: +2 in the >>4 address field
                  if (CG == 1)
                    umma_ss(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0 + b) | k) != 0));
                  else
                    umma_ss_cg2(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (uint32_t)(((kb - kb0 + b) | k) != 0));
                }
                if (CG == 1) umma_commit(&empty_bar[st[b]]);
                else umma_commit_cg2(&empty_bar[st[b]]);
              }
            }
          }
          __syncwarp();
          for (int b = 0; b < nb; ++b)
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1;
            }
        }
        if (elect_one()) {
          if (CG == 1) umma_commit(&tfull_bar[as]);
          else umma_commit_cg2(&tfull_bar[as]);
        }
        __syncwarp();
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 144;");
    // ================================ epilogue ====================================
    const int q = warp & 3;  // TMEM lane quarter owned by this warp
    const int sub = (warp - kFirstEpiWarp) >> 2;  // which of the 3 chunk groups this warp serves
    const int r = q * 32 + lane;
    const ctrlv_epilogue& ep = p.ep;
    const int n_out_total = ep.geglu ? p.N / 2 : p.N;
    const int n_store = ep.n_store > 0 ? ep.n_store : n_out_total;
    const uint32_t tempty_lead0 = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
    const uint32_t tempty_lead1 = (CG == 2) ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0u;
    // row-owner stores / residual loads (two 256-bit accesses per 64-byte row piece) when nothing needs the
    // staged, transposed tile (no GroupNorm statistics) and every row piece is 32-byte aligned
    const bool rowown = CTRLV_EPI_ROWOWN && !ep.geglu && ep.gn_sums == nullptr && ep.out != nullptr && (ep.ld_out % 16) == 0 &&
                        (reinterpret_cast<uintptr_t>(ep.out) & 31) == 0 &&
                        (ep.res1 == nullptr || ((ep.ld_res1 % 16) == 0 && (reinterpret_cast<uintptr_t>(ep.res1) & 31) == 0));
    // position of this thread's row inside the tile box (loop invariant)
    const int ix = r % p.bx;
    const int iy = (r / p.bx) % p.by;
    const int iz = r / (p.bx * p.by);
    WorkIter<SK> wi;
    wi.init(p, unit0, nunits);
    int tile, kb0, kb1;
    for (int it = 0; wi.next(p, tile, kb0, kb1); ++it) {
      uint32_t q_, nt_, tx_, ty_, tz_;
      fd_divmod((uint32_t)tile, p.fd_n, q_, nt_);
      fd_divmod(q_ * CG + crank, p.fd_x, q_, tx_);
      fd_divmod(q_, p.fd_y, tz_, ty_);
      const int nt = (int)nt_, tx = (int)tx_, ty = (int)ty_, tz = (int)tz_;
      const int x = tx * p.bx + ix, y = ty * p.by + iy, z = tz * p.bz + iz;
      const bool valid = (iz < p.bz) && (x < p.X) && (y < p.Y) && (z < p.Z);
      const long long m = ((long long)z * p.oY + y * p.omy + p.ooy) * p.oX + x * p.omx + p.oox;

      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;

      // rowbias: if every valid row of this warp uses the same table row it is folded into the bias
      // fetch (one value per lane); otherwise each thread reads its own row in ep_chunk
      const float* rb = nullptr;
      const float* rb_uniform = nullptr;
      if (ep.rb_mode != 0) {
        const int my_ridx = valid ? rowbias_index(ep, (int)m) : -1;
        int ref = my_ridx;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ref = max(ref, __shfl_xor_sync(0xffffffffu, ref, o));
        const bool uni = __all_sync(0xffffffffu, my_ridx < 0 || my_ridx == ref);
        if (uni && ref >= 0) rb_uniform = ep.rowbias + (size_t)ref * ep.ld_rowbias;
        else if (valid) rb = ep.rowbias + (size_t)my_ridx * ep.ld_rowbias;
      }
      float* sbias = wbias[warp - kFirstEpiWarp];
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.BN);
      uint4* wst = wstage[warp - kFirstEpiWarp];
      if (SK && (kb0 != 0 || kb1 != p.kblocks)) {
        // ---- a k-range of a split tile: park the fp32 partial; igemm_fixup_kernel sums the parts and finishes ----
        const int first_tile = (int)fd_div((uint32_t)(unit0 * p.sk_per), p.fd_kb);
        const size_t slot_elems = (size_t)kBM * p.BN;
        float* mine = p.sk_part + (size_t)((unit0 * 2 + (tile == first_tile ? 0 : 1)) * CG + (int)crank) * slot_elems;
        mbar_wait(&tfull_bar[as], aphase);
        tc_fence_after();
        for (int c = sub; c < p.BN / 32; c += kEpiWarps / 4) {
          uint32_t raw[32];
          tmem_ld32(t_row + (uint32_t)(c * 32), raw);
          tmem_ld_wait();
          float* dst = mine + (size_t)(c * kBM + r) * 32;  // 128 bytes per lane: four whole-sector 256-bit stores
#pragma unroll
          for (int h = 0; h < 4; ++h)
            st_global_v8(dst + 8 * h, raw[8 * h], raw[8 * h + 1], raw[8 * h + 2], raw[8 * h + 3], raw[8 * h + 4],
                         raw[8 * h + 5], raw[8 * h + 6], raw[8 * h + 7]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 1) mbar_arrive(&tempty_bar[as]);
          else mbar_arrive_cluster(as ? tempty_lead1 : tempty_lead0);
        }
        continue;
      }
      if (!SK && ep.geglu)  // (the stream-K schedule is never chosen for a GEGLU epilogue)
        ep_tile<true>(ep, &tfull_bar[as], aphase, t_row, m, nt * p.BN, p.N, p.BN, n_store, valid, sub, sbias, wst, rb, rb_uniform, p, false);
      else
        ep_tile<false>(ep, &tfull_bar[as], aphase, t_row, m, nt * p.BN, p.N, p.BN, n_store, valid, sub, sbias, wst, rb, rb_uniform, p, rowown);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&tempty_bar[as]);
        else mbar_arrive_cluster(as ? tempty_lead1 : tempty_lead0);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer may still read our smem / arrive on our barriers
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    if (CG == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc_cg2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// Stream-K fix-up, launched behind igemm_kernel<CG, true>: one 128-thread block per (split tile, CTA of the pair,
// 32-column chunk), thread = accumulator row as in the epilogue warps.  It adds the tile's fp32 partials in
// contributor order (no atomics: the sum is the same from run to run) and runs the ordinary fused epilogue on it.
// Whole tiles were finished by the GEMM kernel's own epilogue: their blocks exit at once.
__global__ void __launch_bounds__(128) igemm_fixup_kernel(const __grid_constant__ IgemmParams p) {
  __shared__ __align__(16) float wbias[4][32];
  __shared__ __align__(16) uint4 wstage[4][128];
  pdl_trigger();
  const int nch = p.BN / 32;
  int b = (int)blockIdx.x;
  const int c = b % nch; b /= nch;
  const int crank = b % p.cg;
  const int tile = b / p.cg;
  SkParts sk;
  sk.gf = (int)fd_div((uint32_t)(tile * p.kblocks), p.fd_per);
  sk.gl = (int)fd_div((uint32_t)((tile + 1) * p.kblocks - 1), p.fd_per);
  // (every block waits for the GEMM kernel, also the ones with nothing to do: this grid must not complete — and
  // release ITS dependents — before the whole tiles written by the GEMM kernel's own epilogue are complete)
  if (sk.gf == sk.gl) {
    pdl_wait();
    return;
  }
  sk.base = p.sk_part;
  sk.slot_elems = (size_t)kBM * p.BN;
  sk.w_first = ((int)fd_div((uint32_t)(sk.gf * p.sk_per), p.fd_kb) == tile) ? 0 : 1;
  sk.cg = p.cg; sk.crank = crank;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = threadIdx.x;
  const ctrlv_epilogue& ep = p.ep;
  const int n_store = ep.n_store > 0 ? ep.n_store : p.N;
  const bool rowown = CTRLV_EPI_ROWOWN && ep.gn_sums == nullptr && ep.out != nullptr && (ep.ld_out % 16) == 0 &&
                      (reinterpret_cast<uintptr_t>(ep.out) & 31) == 0 &&
                      (ep.res1 == nullptr || ((ep.ld_res1 % 16) == 0 && (reinterpret_cast<uintptr_t>(ep.res1) & 31) == 0));
  uint32_t q_, nt_, tx_, ty_, tz_;
  fd_divmod((uint32_t)tile, p.fd_n, q_, nt_);
  fd_divmod(q_ * (uint32_t)p.cg + (uint32_t)crank, p.fd_x, q_, tx_);
  fd_divmod(q_, p.fd_y, tz_, ty_);
  const int ix = r % p.bx, iy = (r / p.bx) % p.by, iz = r / (p.bx * p.by);
  const int x = (int)tx_ * p.bx + ix, y = (int)ty_ * p.by + iy, z = (int)tz_ * p.bz + iz;
  const bool valid = (iz < p.bz) && (x < p.X) && (y < p.Y) && (z < p.Z);
  const long long m = ((long long)z * p.oY + y * p.omy + p.ooy) * p.oX + x * p.omx + p.oox;
  const int n0 = (int)nt_ * p.BN + c * 32;
  const float* rb = nullptr;
  const float* rb_uniform = nullptr;
  if (ep.rb_mode != 0) {
    const int my_ridx = valid ? rowbias_index(ep, (int)m) : -1;
    int ref = my_ridx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ref = max(ref, __shfl_xor_sync(0xffffffffu, ref, o));
    const bool uni = __all_sync(0xffffffffu, my_ridx < 0 || my_ridx == ref);
    if (uni && ref >= 0) rb_uniform = ep.rowbias + (size_t)ref * ep.ld_rowbias;
    else if (valid) rb = ep.rowbias + (size_t)my_ridx * ep.ld_rowbias;
  }
  pdl_wait();  // the partial tiles (and the residual rows / row-bias table of chained launches) are complete from here
  float bv = 0.f;  // lane j: bias (+ warp-uniform rowbias) of column n0 + j
  if (n0 + lane < p.N) {
    if (ep.bias) bv = __ldg(ep.bias + n0 + lane);
    if (rb_uniform) bv += __ldg(rb_uniform + n0 + lane);
  }
  EpRows<32> rows;
  rows.init(m, valid);
  GnTile gn;
  gn.u0 = -1; gn.multi = false;
  if (ep.gn_sums) gn.init(p, m, valid);
  ResPrefetch<32> pf;
  pf.issue(ep, rows, n0, n_store, n0, p.N, rowown, m, valid);
  __syncwarp();
  ep_chunk<false, true>(ep, 0u, m, valid, n0, n_store, wbias[warp], wstage[warp], rb, bv, pf, rows, gn, p, rowown, &sk,
                        (c * kBM + r) * 32);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
// per-device launch limits (one process may drive several devices: nothing is cached across devices)
struct DevProps { int sms; int max_smem; };
static DevProps g_props[64];  // indexed by device ordinal; sms == 0: not initialised yet

static int device_props(const DevProps** out) {
  int dev = 0;
  CTRLV_CUDA(cudaGetDevice(&dev));
  CTRLV_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  DevProps& dp = g_props[dev];
  if (dp.sms == 0) {
    int sms = 0, smem = 0;
    CTRLV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTRLV_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fa;
    CTRLV_CUDA(cudaFuncGetAttributes(&fa, igemm_kernel<1, false>));
    smem -= (int)fa.sharedSizeBytes;  // static shared memory (barriers, epilogue tiles) counts against the limit
    CTRLV_CUDA(cudaFuncSetAttribute(igemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CTRLV_CUDA(cudaFuncSetAttribute(igemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CTRLV_CUDA(cudaFuncSetAttribute(igemm_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CTRLV_CUDA(cudaFuncSetAttribute(igemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dp.max_smem = smem;
    dp.sms = sms;
  }
  *out = &dp;
  return CTRLV_OK;
}

// tile-plan overrides for tuning sweeps (ctrlv_igemm_override; 0 = heuristic)
static int g_force_bn = 0, g_force_cg = 0, g_force_stages = 0;
static int g_streamk_mode = 0;  // ctrlv_igemm_streamk: 0 = heuristic, 1 = never, 2 = whenever the problem allows it

constexpr int kSkMinKblocks = 16;           // shorter K loops are epilogue-paced: nothing to balance

// Stream-K is worth its fix-up launch only where whole tiles fill the SMs badly AND the K loop is very long.
// Measured on a B200 (profiles/r02_streamk_microbench.json, cold L2): conv3x3 5x8 2560 -> 1280 (K = 23040, 10 row
// tiles) 90.8 -> 78.7 us; conv3x3 1280 -> 1280 (K = 11520) 52.1 -> 51.2; conv(3,1,1) (K = 3840) 25.4 -> 35.6; Linear
// 1280 -> 1280 at M = 1120 16.4 -> 29.7.  These problems are not limited by idle SMs but by the L2 -> SM operand
// stream (about 6300 B/clk for the whole chip): cutting K spreads the same bytes over more SMs and adds the partial
// tiles and a second launch.  `slots` = CTAs (pairs) that run concurrently.
static bool sk_wanted(int tiles_total, int slots, int kblocks) {
  if (g_streamk_mode == 1 || kblocks < kSkMinKblocks) return false;
  if (g_streamk_mode == 2) return true;
  const int waves = (tiles_total + slots - 1) / slots;
  return kblocks >= 256 && (double)tiles_total < 0.8 * (double)waves * slots && waves <= 2;
}

// choose the (bx, by, bz) row box (<= 128 rows) that wastes the fewest MMA rows
static void choose_box(int X, int Y, int Z, int* bx, int* by, int* bz) {
  double best = -1.0;
  int bbx = 1, bby = 1, bbz = 1;
  for (int cx = 1; cx <= (X < 128 ? X : 128); ++cx) {
    if (!((X % cx == 0) || (cx == 128))) continue;
    if (cx > 256) break;
    for (int cy = 1; cy <= Y && cx * cy <= 128; ++cy) {
      for (int cz = 1; cz <= Z && cx * cy * cz <= 128; ++cz) {
        const long long tiles = (long long)((X + cx - 1) / cx) * ((Y + cy - 1) / cy) * ((Z + cz - 1) / cz);
        const double eff = (double)X * Y * Z / ((double)tiles * 128.0);
        // tie-break: prefer wide x (long contiguous TMA rows), then y
        const double score = eff + 1e-6 * cx + 1e-9 * cy;
        if (score > best) {
          best = score;
          bbx = cx; bby = cy; bbz = cz;
        }
      }
    }
  }
  *bx = bbx; *by = bby; *bz = bbz;
}

// n-tile: among the tile widths that divide N pick the one with the lowest modelled time
// (waves x per-tile MMA time; narrow tiles are shared-memory bound, so they cost at least 96)
static int choose_bn(int N, long long tiles_m, int num_sms) {
  const int cands[] = {256, 160, 128, 192, 96, 64};
  int best = 0;
  double best_cost = 1e30;
  for (int bn : cands) {
    if (N % bn != 0) continue;
    const long long tiles = tiles_m * (N / bn);
    const long long waves = (tiles + num_sms - 1) / num_sms;
    // a tile costs its width plus a fixed share (pipeline fill/drain, epilogue set-up) worth ~96
    // columns: fitted to the measured n-tile sweep (scripts/sweep_tiles.py)
    const double cost = (double)waves * (bn + 96) + 1e-3 * (256 - bn);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// CTA pairs (cta_group::2) or single CTAs for a problem (see the comment at the call site)
static int choose_cg(int BN, int N, int tiles_m, int kblocks) {
  // refit on the round-2 sweep (profiles/r02_igemm_tile_sweep.json, after the MMA warp started issuing several
  // k-blocks per round): pairs also pay for the level-3 problems (10 m-tiles) from K = 1280 on and for the
  // K = 640 projections from N = 1920 on
  return (BN % 32 == 0 && N % BN == 0 &&
          ((tiles_m >= 16 && (kblocks >= 20 || (kblocks >= 10 && N >= 1920))) ||
           (tiles_m >= 8 && kblocks >= 20 && (N >= 2560 || kblocks >= 60))))
             ? 2 : 1;
}

static int igemm_launch(const ctrlv_igemm_desc* d, cudaStream_t stream) {
  const DevProps* dp = nullptr;
  int rc = device_props(&dp);
  if (rc) return rc;
  const int g_num_sms = dp->sms, g_max_smem = dp->max_smem;
  CTRLV_CHECK_ARG(d != nullptr, "igemm: null descriptor");
  CTRLV_CHECK_ARG(d->nsrc >= 1 && d->nsrc <= CTRLV_MAX_SRC, "igemm: nsrc=%d out of range", d->nsrc);
  CTRLV_CHECK_ARG(d->nseg >= 1 && d->nseg <= CTRLV_MAX_SEG, "igemm: nseg=%d out of range", d->nseg);
  CTRLV_CHECK_ARG(d->X > 0 && d->Y > 0 && d->Z > 0, "igemm: empty row space %dx%dx%d", d->X, d->Y, d->Z);
  CTRLV_CHECK_ARG(d->N > 0 && d->N % 32 == 0, "igemm: N=%d must be a positive multiple of 32", d->N);
  CTRLV_CHECK_ARG(d->W != nullptr, "igemm: null weights");
  CTRLV_CHECK_ARG((reinterpret_cast<uintptr_t>(d->W) & 15) == 0, "igemm: weights not 16B aligned");

  IgemmParams p;
  memset(&p, 0, sizeof(p));
  p.X = d->X; p.Y = d->Y; p.Z = d->Z;
  choose_box(d->X, d->Y, d->Z, &p.bx, &p.by, &p.bz);
  p.tiles_x = (d->X + p.bx - 1) / p.bx;
  p.tiles_y = (d->Y + p.by - 1) / p.by;
  p.tiles_z = (d->Z + p.bz - 1) / p.bz;
  int kblocks = 0;
  for (int s = 0; s < d->nseg; ++s) {
    const ctrlv_seg& sg = d->seg[s];
    CTRLV_CHECK_ARG(sg.src >= 0 && sg.src < d->nsrc, "igemm: seg %d source %d out of range", s, sg.src);
    CTRLV_CHECK_ARG(sg.nchunk > 0 && sg.c0 % 64 == 0 && sg.c0 + sg.nchunk * 64 <= d->src[sg.src].C,
                    "igemm: seg %d channel range [%d,+%d*64) outside source C=%d", s, sg.c0, sg.nchunk,
                    d->src[sg.src].C);
    p.seg[s].map = sg.src; p.seg[s].c0 = sg.c0; p.seg[s].nchunk = sg.nchunk;
    p.seg[s].dx = sg.dx; p.seg[s].dy = sg.dy; p.seg[s].dz = sg.dz;
    kblocks += sg.nchunk;
  }
  CTRLV_CHECK_ARG(kblocks * 64 == d->K, "igemm: K=%d does not match segments (%d)", d->K, kblocks * 64);
  p.nseg = d->nseg;
  p.kblocks = kblocks;
  const int kblocks_pre = kblocks;
  p.N = d->N;
  p.BN = d->bn > 0 ? d->bn : choose_bn(d->N, (long long)p.tiles_x * p.tiles_y * p.tiles_z, g_num_sms);
  if (g_force_bn) p.BN = g_force_bn;
  if (p.BN == 0) {
    // ragged N: largest 32-multiple tile, TMA zero-fills the weight rows past N
    p.BN = d->N >= 256 ? 256 : ((d->N + 31) / 32) * 32;
  }
  CTRLV_CHECK_ARG(p.BN % 32 == 0 && p.BN >= 32 && p.BN <= 256, "igemm: bad n-tile %d", p.BN);
  p.tiles_n = (d->N + p.BN - 1) / p.BN;
  const int tiles_m = p.tiles_x * p.tiles_y * p.tiles_z;
  // CTA pairs (cta_group::2): a single CTA's 128 x 256 tile is shared-memory-bandwidth bound
  // (TMA writes + MMA reads ~ 96 KB per k-block); a pair halves the B traffic and deepens the TMA
  // ring.  Measured on this path's shapes (profiles/r01_cg2_vs_cg1.txt): +15-19 % on the 3x3 convs,
  // a win for K >= 1280 and for wide-N K = 640, a loss for K = 320 and tiny M.
  p.cg = choose_cg(p.BN, d->N, tiles_m, kblocks_pre);
  if (g_force_cg) p.cg = g_force_cg;
  p.tiles_total = ((tiles_m + p.cg - 1) / p.cg) * p.tiles_n;
  // Stream-K (needs the caller's workspace, ctrlv_epilogue.splitk_ws): with every SM busy for the same number of
  // k-blocks whatever the tile count, the widest tile is the cheapest — 256 columns, CTA pairs
  bool sk = false;
  if (d->ep.splitk_ws != nullptr && !d->ep.geglu && d->N % p.BN == 0 &&
      sk_wanted(p.tiles_total, g_num_sms / p.cg, kblocks_pre)) {
    int bn = p.BN, cg = p.cg;
    if (!g_force_bn && d->bn <= 0) {
      const int wide[] = {256, 192, 160, 128};
      for (int w : wide)
        if (d->N % w == 0) { bn = w; break; }
    }
    if (!g_force_cg && tiles_m >= 2) cg = 2;
    const int tiles_n = d->N / bn;
    const long long tt = (long long)((tiles_m + cg - 1) / cg) * tiles_n;
    const long long total = tt * kblocks_pre;
    int G = g_num_sms / cg;
    if (total < (long long)G * 8) G = (int)(total / 8);  // at least 8 k-blocks per CTA
    {  // a tile of K k-blocks starting anywhere inside a range of `per` touches <= ceil((per - 1 + K) / per) ranges
      const long long per_min = (kblocks_pre - 1 + kSkMaxContrib - 2) / (kSkMaxContrib - 1);
      if (G >= 1 && (total + G - 1) / G < per_min) G = (int)(total / per_min);
    }
    const long long need = (long long)G * cg * 2 * kBM * bn * (long long)sizeof(float);
    if (G >= 2 && need <= d->ep.splitk_bytes && total < (1ll << 30) && tt * cg * (bn / 32) < (1ll << 30) &&
        (reinterpret_cast<uintptr_t>(d->ep.splitk_ws) & 31) == 0) {
      sk = true;
      p.BN = bn; p.cg = cg;
      p.tiles_n = tiles_n;
      p.tiles_total = (int)tt;
      p.sk_per = (int)((total + G - 1) / G);
      p.sk_total = (int)total;
      p.fd_kb = make_fastdiv((uint32_t)kblocks_pre);
      p.fd_per = make_fastdiv((uint32_t)p.sk_per);
      p.sk_part = reinterpret_cast<float*>(d->ep.splitk_ws);
    }
  }
  if (d->out_X > 0) {
    CTRLV_CHECK_ARG(d->out_Y > 0 && d->out_mul_x >= 1 && d->out_mul_y >= 1 && d->out_off_x >= 0 && d->out_off_y >= 0 &&
                        (d->X - 1) * d->out_mul_x + d->out_off_x < d->out_X && (d->Y - 1) * d->out_mul_y + d->out_off_y < d->out_Y,
                    "igemm: output-row remap does not fit the %dx%d output grid", d->out_X, d->out_Y);
    p.omx = d->out_mul_x; p.omy = d->out_mul_y; p.oox = d->out_off_x; p.ooy = d->out_off_y; p.oX = d->out_X; p.oY = d->out_Y;
  } else {
    p.omx = 1; p.omy = 1; p.oox = 0; p.ooy = 0; p.oX = d->X; p.oY = d->Y;
  }
  p.fd_n = make_fastdiv((uint32_t)p.tiles_n);
  p.fd_x = make_fastdiv((uint32_t)p.tiles_x);
  p.fd_y = make_fastdiv((uint32_t)p.tiles_y);


  for (int i = 0; i < d->nsrc; ++i) {
    const ctrlv_src& s = d->src[i];
    CTRLV_CHECK_ARG(s.ptr != nullptr && (reinterpret_cast<uintptr_t>(s.ptr) & 15) == 0,
                    "igemm: source %d null or not 16B aligned", i);
    CTRLV_CHECK_ARG(s.C % 64 == 0 && s.sx % 8 == 0 && s.sy % 8 == 0 && s.sz % 8 == 0,
                    "igemm: source %d needs C%%64==0 and strides %%8==0", i);
    uint64_t dims[4] = {(uint64_t)s.C, (uint64_t)d->X, (uint64_t)d->Y, (uint64_t)d->Z};
    uint64_t strides[3] = {(uint64_t)s.sx * 2, (uint64_t)s.sy * 2, (uint64_t)s.sz * 2};
    uint32_t box[4] = {64, (uint32_t)p.bx, (uint32_t)p.by, (uint32_t)p.bz};
    rc = encode_tmap_bf16(&p.tmA[i], s.ptr, 4, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    uint64_t strides[1] = {(uint64_t)d->K * 2};
    uint32_t box[2] = {64, (uint32_t)(p.BN / p.cg)};
    rc = encode_tmap_bf16(&p.tmB, d->W, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  p.a_bytes = 64 * p.bx * p.by * p.bz * 2;
  p.stage_bytes = kBM * kBK * 2 + (p.BN / p.cg) * kBK * 2;
  p.m_tiles = tiles_m;
  int stages = (g_max_smem - 2048) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (g_force_stages) stages = g_force_stages;
  CTRLV_CHECK_ARG(stages >= 2, "igemm: not enough shared memory for 2 stages");
  p.stages = stages;
  // k-blocks per round of the MMA warp: 2 halves its per-round overhead everywhere; 4 pays for the long-K convs
  // when the ring is deep enough that waiting for four full stages does not starve the producer (measured on the
  // step's problems: conv3x3 levels 1-3 another 1-5 %, the K <= 2560 Linears lose with 4)
  p.mma_batch = (kblocks >= 40 && stages >= 6) ? 4 : 2;
  if (p.mma_batch > stages) p.mma_batch = stages;
  if (p.mma_batch > kMmaBatch) p.mma_batch = kMmaBatch;
  p.tmem_cols = 2 * p.BN <= 128 ? 128 : (2 * p.BN <= 256 ? 256 : 512);
  p.ep = d->ep;  // s_acc is taken literally: 0 gives out = s_res1*res1 + s_res2*res2 (conditioning_scale = 0)
  CTRLV_CHECK_ARG(p.ep.out != nullptr || p.ep.out_f32 != nullptr, "igemm: no output");
  if (p.ep.geglu) CTRLV_CHECK_ARG(d->N % 64 == 0, "igemm: GEGLU needs N %% 64 == 0");
  if (p.ep.rb_mode != 0)
    CTRLV_CHECK_ARG(p.ep.rowbias != nullptr && p.ep.rb_div > 0, "igemm: rowbias mode without table");
  if (p.ep.rb_mode == 2 || p.ep.rb_mode == 3)
    CTRLV_CHECK_ARG(p.ep.rb_mod > 0 && (p.ep.rb_mode == 2 || p.ep.rb_B > 0) && p.ep.rb_off >= 0, "igemm: bad rowbias modulus");
  if (p.ep.gn_sums) {
    CTRLV_CHECK_ARG(!p.ep.geglu && p.ep.out != nullptr && p.ep.n_store == 0 && d->N % p.BN == 0,
                    "igemm: GroupNorm statistics need a full-width bf16 output without GEGLU");
    // an aligned 8-column piece must fall into at most two groups
    CTRLV_CHECK_ARG(p.ep.gn_rows_per_unit > 0 && (p.ep.gn_cg >= 7 || p.ep.gn_cg == 4 || p.ep.gn_cg == 6) &&
                        p.ep.gn_c_off >= 0 && p.ep.gn_c_off % 8 == 0 && (reinterpret_cast<uintptr_t>(p.ep.gn_sums) & 7) == 0,
                    "igemm: GroupNorm statistics need rows_per_unit > 0, 4, 6 or >= 7 channels per group, c_off %% 8 == 0");
    CTRLV_CHECK_ARG((p.ep.gn_c_off + d->N + p.ep.gn_cg - 1) / p.ep.gn_cg <= 32, "igemm: channels fall outside the 32 groups");
    CTRLV_CHECK_ARG(p.ep.gn_rep >= 1 && (p.ep.gn_rep & (p.ep.gn_rep - 1)) == 0 && p.ep.gn_units >= 1,
                    "igemm: gn_rep must be a power of two, gn_units >= 1");
    p.fd_rpu = make_fastdiv((uint32_t)p.ep.gn_rows_per_unit);
    p.fd_cg = make_fastdiv((uint32_t)p.ep.gn_cg);
  }
  // vector paths need 16B-aligned rows
  const int n_out = p.ep.geglu ? d->N / 2 : d->N;
  const int n_store = p.ep.n_store > 0 ? p.ep.n_store : n_out;
  if (n_store >= 16) {
    if (p.ep.out) CTRLV_CHECK_ARG(p.ep.ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(p.ep.out) & 15) == 0, "igemm: out must be 16B aligned (ld %% 8)");
    if (p.ep.res1) CTRLV_CHECK_ARG(p.ep.ld_res1 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.ep.res1) & 15) == 0, "igemm: res1 must be 16B aligned");
    if (p.ep.res2) CTRLV_CHECK_ARG(p.ep.ld_res2 % 8 == 0 && (reinterpret_cast<uintptr_t>(p.ep.res2) & 15) == 0, "igemm: res2 must be 16B aligned");
    if (p.ep.out_f32) CTRLV_CHECK_ARG(p.ep.ld_out_f32 % 4 == 0, "igemm: out_f32 ld %% 4");
  }

  size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const int threads = kThreads;
  if (sk) {
    const int G = (p.sk_total + p.sk_per - 1) / p.sk_per;  // CTAs (pairs) with a non-empty k-block range
    if (p.cg == 1) CTRLV_CUDA(launch_pdl(igemm_kernel<1, true>, dim3(G), dim3(threads), smem, stream, p));
    else CTRLV_CUDA(launch_cluster2(igemm_kernel<2, true>, dim3(2 * G), dim3(threads), smem, stream, p));
    CTRLV_CUDA(launch_pdl(igemm_fixup_kernel, dim3(p.tiles_total * p.cg * (p.BN / 32)), dim3(128), (size_t)0, stream, p));
  } else if (p.cg == 1) {
    int grid = p.tiles_total < g_num_sms ? p.tiles_total : g_num_sms;
    CTRLV_CUDA(launch_pdl(igemm_kernel<1, false>, dim3(grid), dim3(threads), smem, stream, p));
  } else {
    const int pairs = p.tiles_total < g_num_sms / 2 ? p.tiles_total : g_num_sms / 2;
    CTRLV_CUDA(launch_cluster2(igemm_kernel<2, false>, dim3(2 * pairs), dim3(threads), smem, stream, p));
  }
  return CTRLV_OK;
}

}  // namespace ctrlv

using namespace ctrlv;

extern "C" int ctrlv_igemm_override(int32_t bn, int32_t cta_group, int32_t stages) {
  CTRLV_CHECK_ARG((bn == 0 || (bn % 32 == 0 && bn >= 32 && bn <= 256)) && cta_group >= 0 && cta_group <= 2 &&
                      stages >= 0 && stages <= kMaxStages && stages != 1,
                  "igemm_override: bn=%d cta_group=%d stages=%d", bn, cta_group, stages);
  g_force_bn = bn; g_force_cg = cta_group; g_force_stages = stages;
  return CTRLV_OK;
}

extern "C" int ctrlv_igemm_streamk(int32_t mode) {
  CTRLV_CHECK_ARG(mode >= 0 && mode <= 2, "igemm_streamk: mode %d (0 = heuristic, 1 = never, 2 = whenever possible)", mode);
  g_streamk_mode = mode;
  return CTRLV_OK;
}

extern "C" int ctrlv_igemm(const ctrlv_igemm_desc* desc, void* stream) {
  return igemm_launch(desc, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ctrlv_linear(const void* A, int64_t lda, int32_t M, int32_t K, const void* W,
                            int32_t N, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr, "linear: null epilogue");
  CTRLV_CHECK_ARG(K % 64 == 0, "linear: K=%d must be a multiple of 64", K);
  ctrlv_igemm_desc d;
  memset(&d, 0, sizeof(d));
  d.nsrc = 1;
  d.src[0].ptr = A; d.src[0].C = K; d.src[0].sx = lda; d.src[0].sy = lda * (int64_t)M; d.src[0].sz = lda * (int64_t)M;
  d.X = M; d.Y = 1; d.Z = 1;
  d.nseg = 1;
  d.seg[0].src = 0; d.seg[0].c0 = 0; d.seg[0].nchunk = K / 64;
  d.W = W; d.N = N; d.K = K;
  d.ep = *ep;
  return igemm_launch(&d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ctrlv_conv3x3(const void* src0, int32_t C0, const void* src1, int32_t C1,
                             int32_t frames, int32_t H, int32_t Wd, int32_t stride,
                             const void* sc0, int32_t SC0, const void* sc1, int32_t SC1,
                             const void* W, int32_t N, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr, "conv3x3: null epilogue");
  CTRLV_CHECK_ARG(stride == 1 || stride == 2, "conv3x3: stride %d unsupported", stride);
  CTRLV_CHECK_ARG(src0 != nullptr && C0 > 0 && C0 % 64 == 0 && C1 % 64 == 0 && SC0 % 64 == 0 && SC1 % 64 == 0,
                  "conv3x3: channel counts must be multiples of 64");
  const int Cin = C0 + (src1 ? C1 : 0);
  ctrlv_igemm_desc d;
  memset(&d, 0, sizeof(d));
  int nseg = 0;
  if (stride == 1) {
    d.X = Wd; d.Y = H; d.Z = frames;
    d.nsrc = 0;
    const int i0 = d.nsrc++;
    d.src[i0].ptr = src0; d.src[i0].C = C0; d.src[i0].sx = C0; d.src[i0].sy = (int64_t)C0 * Wd; d.src[i0].sz = (int64_t)C0 * Wd * H;
    int i1 = -1;
    if (src1) {
      i1 = d.nsrc++;
      d.src[i1].ptr = src1; d.src[i1].C = C1; d.src[i1].sx = C1; d.src[i1].sy = (int64_t)C1 * Wd; d.src[i1].sz = (int64_t)C1 * Wd * H;
    }
    // K order: [tap][src0 channels | src1 channels], then the optional raw shortcut sources
    for (int t = 0; t < 9; ++t) {
      ctrlv_seg& s = d.seg[nseg++];
      s.src = i0; s.c0 = 0; s.nchunk = C0 / 64; s.dx = t % 3 - 1; s.dy = t / 3 - 1; s.dz = 0;
      if (src1) {
        ctrlv_seg& s1 = d.seg[nseg++];
        s1.src = i1; s1.c0 = 0; s1.nchunk = C1 / 64; s1.dx = t % 3 - 1; s1.dy = t / 3 - 1; s1.dz = 0;
      }
    }
    if (sc0) {
      const int j0 = d.nsrc++;
      d.src[j0].ptr = sc0; d.src[j0].C = SC0; d.src[j0].sx = SC0; d.src[j0].sy = (int64_t)SC0 * Wd; d.src[j0].sz = (int64_t)SC0 * Wd * H;
      ctrlv_seg& s = d.seg[nseg++];
      s.src = j0; s.c0 = 0; s.nchunk = SC0 / 64; s.dx = s.dy = s.dz = 0;
      if (sc1) {
        const int j1 = d.nsrc++;
        d.src[j1].ptr = sc1; d.src[j1].C = SC1; d.src[j1].sx = SC1; d.src[j1].sy = (int64_t)SC1 * Wd; d.src[j1].sz = (int64_t)SC1 * Wd * H;
        ctrlv_seg& s2 = d.seg[nseg++];
        s2.src = j1; s2.c0 = 0; s2.nchunk = SC1 / 64; s2.dx = s2.dy = s2.dz = 0;
      }
    }
    d.K = 9 * Cin + (sc0 ? SC0 + (sc1 ? SC1 : 0) : 0);
  } else {
    CTRLV_CHECK_ARG(src1 == nullptr && sc0 == nullptr, "conv3x3: stride 2 takes one source, no shortcut");
    CTRLV_CHECK_ARG(H % 2 == 0 && Wd % 2 == 0, "conv3x3: stride 2 needs even H, W");
    const int Ho = H / 2, Wo = Wd / 2;
    d.X = Wo; d.Y = Ho; d.Z = frames;
    // four parity sub-lattices of the input: (py, px); input pixel (2*oy + dy - 1, 2*ox + dx - 1)
    d.nsrc = 4;
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        ctrlv_src& s = d.src[py * 2 + px];
        s.ptr = reinterpret_cast<const bf16*>(src0) + ((int64_t)py * Wd + px) * C0;
        s.C = C0; s.sx = 2 * (int64_t)C0; s.sy = 2 * (int64_t)C0 * Wd; s.sz = (int64_t)C0 * Wd * H;
      }
    for (int t = 0; t < 9; ++t) {
      const int dy = t / 3 - 1, dx = t % 3 - 1;
      const int py = dy & 1, px = dx & 1;          // -1 -> 1, 0 -> 0, 1 -> 1
      const int oy = (dy - py) / 2, ox = (dx - px) / 2;  // -1 -> -1, 0 -> 0, 1 -> 0
      ctrlv_seg& s = d.seg[nseg++];
      s.src = py * 2 + px; s.c0 = 0; s.nchunk = C0 / 64; s.dx = ox; s.dy = oy; s.dz = 0;
    }
    d.K = 9 * C0;
  }
  d.nseg = nseg;
  d.W = W; d.N = N;
  d.ep = *ep;
  return igemm_launch(&d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ctrlv_conv3x3_s2_pad01(const void* src, int32_t C, int32_t frames, int32_t H, int32_t Wd,
                                      const void* W, int32_t N, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr && src != nullptr, "conv3x3_s2_pad01: null argument");
  CTRLV_CHECK_ARG(C > 0 && C % 64 == 0, "conv3x3_s2_pad01: C=%d must be a multiple of 64", C);
  CTRLV_CHECK_ARG(H % 2 == 0 && Wd % 2 == 0, "conv3x3_s2_pad01: needs even H, W");
  ctrlv_igemm_desc d;
  memset(&d, 0, sizeof(d));
  d.X = Wd / 2; d.Y = H / 2; d.Z = frames;
  // out(oy, ox) = sum_k w[ky][kx] x(2oy + ky, 2ox + kx): tap k reads parity lattice k & 1 at offset k >> 1;
  // the row / column past the frame (the F.pad(0,1,0,1) of diffusers' Downsample2D(padding=0)) is TMA zero fill
  d.nsrc = 4;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      ctrlv_src& s = d.src[py * 2 + px];
      s.ptr = reinterpret_cast<const bf16*>(src) + ((int64_t)py * Wd + px) * C;
      s.C = C; s.sx = 2 * (int64_t)C; s.sy = 2 * (int64_t)C * Wd; s.sz = (int64_t)C * Wd * H;
    }
  d.nseg = 9;
  for (int t = 0; t < 9; ++t) {
    const int ky = t / 3, kx = t % 3;
    ctrlv_seg& s = d.seg[t];
    s.src = (ky & 1) * 2 + (kx & 1); s.c0 = 0; s.nchunk = C / 64; s.dx = kx >> 1; s.dy = ky >> 1; s.dz = 0;
  }
  d.W = W; d.N = N; d.K = 9 * C;
  d.ep = *ep;
  return igemm_launch(&d, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int ctrlv_upsample2x_conv3x3(const void* src, int32_t C, int32_t frames, int32_t H, int32_t Wd,
                                        const void* Wp, int32_t N, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr && src != nullptr && Wp != nullptr, "upsample2x_conv3x3: null argument");
  CTRLV_CHECK_ARG(C > 0 && C % 64 == 0, "upsample2x_conv3x3: C=%d must be a multiple of 64", C);
  CTRLV_CHECK_ARG(ep->rb_mode == 0 && ep->res1 == nullptr && ep->res2 == nullptr,
                  "upsample2x_conv3x3: bias-only epilogue");
  // conv3x3(nearest2x(x)) at output pixel (2y+py, 2x+px) only sees a 2x2 patch of x: four phase convs with
  // summed taps (phase matrix [N][4*C], patch order (dy, dx) increasing), written to their strided rows
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      ctrlv_igemm_desc d;
      memset(&d, 0, sizeof(d));
      d.nsrc = 1;
      d.src[0].ptr = src; d.src[0].C = C; d.src[0].sx = C; d.src[0].sy = (int64_t)C * Wd; d.src[0].sz = (int64_t)C * Wd * H;
      d.X = Wd; d.Y = H; d.Z = frames;
      d.nseg = 4;
      int i = 0;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          ctrlv_seg& s = d.seg[i++];
          s.src = 0; s.c0 = 0; s.nchunk = C / 64; s.dx = (px == 0 ? -1 : 0) + b; s.dy = (py == 0 ? -1 : 0) + a; s.dz = 0;
        }
      d.W = reinterpret_cast<const bf16*>(Wp) + (size_t)(py * 2 + px) * N * 4 * C;
      d.N = N; d.K = 4 * C;
      d.out_mul_x = 2; d.out_mul_y = 2; d.out_off_x = px; d.out_off_y = py; d.out_X = 2 * Wd; d.out_Y = 2 * H;
      d.ep = *ep;
      const int rc = igemm_launch(&d, reinterpret_cast<cudaStream_t>(stream));
      if (rc) return rc;
    }
  return CTRLV_OK;
}

extern "C" int ctrlv_conv_t3(const void* src, int32_t C, int32_t B, int32_t T, int32_t HW,
                             const void* W, int32_t N, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr, "conv_t3: null epilogue");
  CTRLV_CHECK_ARG(C % 64 == 0, "conv_t3: C=%d must be a multiple of 64", C);
  ctrlv_igemm_desc d;
  memset(&d, 0, sizeof(d));
  d.nsrc = 1;
  d.src[0].ptr = src; d.src[0].C = C; d.src[0].sx = C; d.src[0].sy = (int64_t)C * HW; d.src[0].sz = (int64_t)C * HW * T;
  d.X = HW; d.Y = T; d.Z = B;
  d.nseg = 3;
  for (int t = 0; t < 3; ++t) {
    d.seg[t].src = 0; d.seg[t].c0 = 0; d.seg[t].nchunk = C / 64; d.seg[t].dx = 0; d.seg[t].dy = t - 1; d.seg[t].dz = 0;
  }
  d.W = W; d.N = N; d.K = 3 * C;
  d.ep = *ep;
  return igemm_launch(&d, reinterpret_cast<cudaStream_t>(stream));
}

// Tile plan of a problem without launching it (no CUDA calls: usable on a host without a GPU): the
// row-box, the n-tile width and the cta_group the launcher would pick on a device with `num_sms` SMs.
extern "C" int ctrlv_igemm_plan(const ctrlv_igemm_desc* d, int32_t num_sms, int32_t* box_xyz, int32_t* bn,
                                int32_t* cta_group) {
  CTRLV_CHECK_ARG(d != nullptr && box_xyz != nullptr && bn != nullptr && cta_group != nullptr && num_sms > 0,
                  "igemm_plan: bad arguments");
  CTRLV_CHECK_ARG(d->X > 0 && d->Y > 0 && d->Z > 0 && d->N > 0 && d->N % 32 == 0 && d->nseg >= 1 && d->nseg <= CTRLV_MAX_SEG,
                  "igemm_plan: bad descriptor");
  int bx, by, bz;
  choose_box(d->X, d->Y, d->Z, &bx, &by, &bz);
  const int tiles_m = ((d->X + bx - 1) / bx) * ((d->Y + by - 1) / by) * ((d->Z + bz - 1) / bz);
  int kblocks = 0;
  for (int s = 0; s < d->nseg; ++s) kblocks += d->seg[s].nchunk;
  int BN = d->bn > 0 ? d->bn : choose_bn(d->N, tiles_m, num_sms);
  if (BN == 0) BN = d->N >= 256 ? 256 : ((d->N + 31) / 32) * 32;
  box_xyz[0] = bx; box_xyz[1] = by; box_xyz[2] = bz;
  *bn = BN;
  *cta_group = choose_cg(BN, d->N, tiles_m, kblocks);
  return CTRLV_OK;
}
