// GroupNorm (statistics + apply [+ SiLU], one or two channel-concatenated sources) and LayerNorm
// on channels-last bf16 rows.  HBM-bound kernels: 16-byte vector loads/stores, fp32 statistics,
// deterministic two-stage reduction (no float atomics in global memory).
#include <cstdlib>

#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

constexpr int kGroups = 32;
constexpr int kMaxSplit = 256;

// ------------------------------------------------------------------------------------------
// GroupNorm statistics: partial[unit][split][group][2] = (sum, sum of squares)
// Thread layout: nthreads = vpr * rows_par (vpr = 16B vectors per row); thread t owns vector
// column t % vpr for the whole kernel, so consecutive threads read consecutive 16 B.
// ------------------------------------------------------------------------------------------
__global__ void gn_stats_kernel(const bf16* __restrict__ src0, int C0, const bf16* __restrict__ src1,
                                int C1, int rows_per_unit, int rows_per_split, int nsplit,
                                int nwork, float* __restrict__ partial) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float gn_smem[];  // [nwork][8] thread sums, then [vpr*4][2] channel-pair sums
  const int C = C0 + C1;
  const int vpr = C >> 3;
  const int cg = C / kGroups;
  const int unit = blockIdx.x / nsplit;
  const int split = blockIdx.x % nsplit;
  const int tid = threadIdx.x;
  const int vec = tid % vpr;
  const int rpar = nwork / vpr;
  const int rsub = tid < nwork ? tid / vpr : rows_per_unit;  // padding threads do no rows
  const int c = vec << 3;
  const bf16* base;
  int ld;
  if (c < C0) { base = src0 + c; ld = C0; } else { base = src1 + (c - C0); ld = C1; }
  const int r_begin = split * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > rows_per_unit) r_end = rows_per_unit;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t row0 = (size_t)unit * rows_per_unit;
#pragma unroll 4
  for (int r = r_begin + rsub; r < r_end; r += rpar) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (row0 + r) * ld));
    float2 f;
    f = unpack_bf16x2(u.x); s[0] += f.x + f.y; q[0] += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u.y); s[1] += f.x + f.y; q[1] += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u.z); s[2] += f.x + f.y; q[2] += f.x * f.x + f.y * f.y;
    f = unpack_bf16x2(u.w); s[3] += f.x + f.y; q[3] += f.x * f.x + f.y * f.y;
  }
  // deterministic block reduction (no atomics): thread sums -> channel-pair sums -> group sums
  if (tid < nwork) {
    float* t = gn_smem + (size_t)tid * 8;
    t[0] = s[0]; t[1] = s[1]; t[2] = s[2]; t[3] = s[3];
    t[4] = q[0]; t[5] = q[1]; t[6] = q[2]; t[7] = q[3];
  }
  __syncthreads();
  const int npair = vpr * 4;
  float* pair = gn_smem + (size_t)nwork * 8;  // [npair][2]
  for (int pp = tid; pp < npair; pp += blockDim.x) {
    const int v = pp >> 2, pi = pp & 3;
    float ps = 0.f, pq = 0.f;
    for (int rs = 0; rs < rpar; ++rs) {
      const float* t = gn_smem + (size_t)(rs * vpr + v) * 8;
      ps += t[pi];
      pq += t[4 + pi];
    }
    pair[pp * 2] = ps; pair[pp * 2 + 1] = pq;
  }
  __syncthreads();
  if (tid < kGroups) {
    const int ppg = cg >> 1;  // channel pairs per group
    float gs = 0.f, gq = 0.f;
    for (int i = 0; i < ppg; ++i) {
      gs += pair[(tid * ppg + i) * 2];
      gq += pair[(tid * ppg + i) * 2 + 1];
    }
    float* o = partial + ((size_t)unit * nsplit + split) * (kGroups * 2) + tid * 2;
    o[0] = gs; o[1] = gq;
  }
}

// ------------------------------------------------------------------------------------------
// out = a*x + b*y (the ControlNet residual added into a UNet skip connection) fused with the statistics
// of `out` for the GroupNorm that consumes it: same thread layout and block reduction as
// gn_stats_kernel; the block's group sums are added to sums[unit][group] as int64 fixed point
// (x 2^16, integer atomics: order-independent, hence deterministic) — the format the igemm epilogue uses.
// ------------------------------------------------------------------------------------------
__global__ void axpby_gn_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, float a, float b, int C,
                                int rows_per_unit, int rows_per_split, int nsplit, int nwork,
                                bf16* __restrict__ out, unsigned long long* __restrict__ sums, int cg, int c_off,
                                int n_units, int n_rep) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float gn_smem[];  // [nwork][8] thread sums, then [vpr*4][2] channel-pair sums
  const int vpr = C >> 3;
  const int unit = blockIdx.x / nsplit;
  const int split = blockIdx.x % nsplit;
  const int tid = threadIdx.x;
  const int vec = tid % vpr;
  const int rpar = nwork / vpr;
  const int rsub = tid < nwork ? tid / vpr : rows_per_unit;
  const int r_begin = split * rows_per_split;
  int r_end = r_begin + rows_per_split;
  if (r_end > rows_per_unit) r_end = rows_per_unit;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t row0 = (size_t)unit * rows_per_unit;
#pragma unroll 4
  for (int r = r_begin + rsub; r < r_end; r += rpar) {
    const size_t off = (row0 + r) * C + (vec << 3);
    const uint4 ux = __ldg(reinterpret_cast<const uint4*>(x + off));
    uint32_t wo[4] = {ux.x, ux.y, ux.z, ux.w};
    if (y) {  // (y == nullptr: statistics of x itself, nothing written)
      const uint4 uy = __ldg(reinterpret_cast<const uint4*>(y + off));
      const uint32_t wy[4] = {uy.x, uy.y, uy.z, uy.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 fx = unpack_bf16x2(wo[k]), fy = unpack_bf16x2(wy[k]);
        wo[k] = pack_bf16x2(a * fx.x + b * fy.x, a * fx.y + b * fy.y);
      }
      *reinterpret_cast<uint4*>(out + off) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = unpack_bf16x2(wo[k]);  // statistics of the value the consumer will read
      s[k] += f.x + f.y;
      q[k] += f.x * f.x + f.y * f.y;
    }
  }
  if (tid < nwork) {
    float* t = gn_smem + (size_t)tid * 8;
    t[0] = s[0]; t[1] = s[1]; t[2] = s[2]; t[3] = s[3];
    t[4] = q[0]; t[5] = q[1]; t[6] = q[2]; t[7] = q[3];
  }
  __syncthreads();
  const int npair = vpr * 4;
  float* pair = gn_smem + (size_t)nwork * 8;  // [npair][2]
  for (int pp = tid; pp < npair; pp += blockDim.x) {
    const int v = pp >> 2, pi = pp & 3;
    float ps = 0.f, pq = 0.f;
    for (int rs = 0; rs < rpar; ++rs) {
      const float* t = gn_smem + (size_t)(rs * vpr + v) * 8;
      ps += t[pi];
      pq += t[4 + pi];
    }
    pair[pp * 2] = ps; pair[pp * 2 + 1] = pq;
  }
  __syncthreads();
  if (tid < kGroups) {
    // channel pairs [lo, hi) of group tid that fall inside this tensor's channel range [c_off, c_off + C)
    int lo = tid * cg - c_off, hi = lo + cg;
    if (lo < 0) lo = 0;
    if (hi > C) hi = C;
    if (lo < hi) {
      float gs = 0.f, gq = 0.f;
      for (int i = lo >> 1; i < (hi >> 1); ++i) {
        gs += pair[i * 2];
        gq += pair[i * 2 + 1];
      }
      unsigned long long* o = sums + ((size_t)(blockIdx.x & (unsigned)(n_rep - 1)) * n_units + unit) * (kGroups * 2) + tid * 2;
      atomicAdd(o, (unsigned long long)__float2ll_rn(gs * 65536.0f));
      atomicAdd(o + 1, (unsigned long long)__float2ll_rn(gq * 65536.0f));
    }
  }
}

// ------------------------------------------------------------------------------------------
// GroupNorm apply (+SiLU): out[row][C0+C1] = act((x - mean) * rstd * gamma + beta)
// grid = n_units * blocks_per_unit; a block stays inside one statistics unit.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 3) gn_apply_kernel(const bf16* __restrict__ src0, int C0, const bf16* __restrict__ src1,
                                int C1, int rows_per_unit, int blocks_per_unit, int nsplit,
                                int nwork, const float* __restrict__ partial, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float eps, int silu,
                                bf16* __restrict__ out, const long long* __restrict__ fixed_sums, int n_units,
                                int n_rep) {
  pdl_wait();
  pdl_trigger();
  __shared__ float mean_s[kGroups], rstd_s[kGroups];
  __shared__ double wsum[16][kGroups][2];
  const int C = C0 + C1;
  const int vpr = C >> 3;
  const int cg = C / kGroups;
  const int unit = blockIdx.x / blocks_per_unit;
  const int blk = blockIdx.x % blocks_per_unit;
  const int tid = threadIdx.x;
  if (fixed_sums) {
    // statistics accumulated by the producers of the input (igemm epilogue / axpby_gn): int64 fixed point
    // (x 2^16), n_rep replicas of one 512-byte record per unit.  Warp w adds replicas w, w + nwarps, ...
    // (lane = group, one 16-byte load each), then group g adds the warp sums: integer arithmetic, exact.
    const int nwarps = blockDim.x >> 5;
    const int wid = tid >> 5, lane = tid & 31;
    long long (*wll)[kGroups][2] = reinterpret_cast<long long (*)[kGroups][2]>(wsum);
    const longlong2* rec = reinterpret_cast<const longlong2*>(fixed_sums) + (size_t)unit * kGroups + lane;
    long long fs = 0, fq = 0;
    for (int i = wid; i < n_rep; i += nwarps) {
      const longlong2 v = __ldg(rec + (size_t)i * n_units * kGroups);
      fs += v.x;
      fq += v.y;
    }
    wll[wid][lane][0] = fs;
    wll[wid][lane][1] = fq;
    __syncthreads();
    if (tid < kGroups) {
      fs = 0; fq = 0;
      const int nw = nwarps < n_rep ? nwarps : n_rep;
      for (int w = 0; w < nw; ++w) { fs += wll[w][tid][0]; fq += wll[w][tid][1]; }
      const double cnt = (double)rows_per_unit * cg;
      const double mean = (double)fs * (1.0 / 65536.0) / cnt;
      double var = (double)fq * (1.0 / 65536.0) / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      mean_s[tid] = (float)mean;
      rstd_s[tid] = (float)(1.0 / sqrt(var + (double)eps));
    }
  } else
  // deterministic reduction of the split partials: warp w sums splits w, w+nwarps, ... (lane =
  // group, one coalesced 256-byte record per split, loads unrolled so they overlap), then group g
  // adds the warp sums in warp order (fp64)
  {
    const int nwarps = blockDim.x >> 5;
    const int wid = tid >> 5, lane = tid & 31;
    const float2* rec = reinterpret_cast<const float2*>(partial) + (size_t)unit * nsplit * kGroups + lane;
    double s = 0.0, q = 0.0;
#pragma unroll 4
    for (int i = wid; i < nsplit; i += nwarps) {
      const float2 v = __ldg(rec + (size_t)i * kGroups);
      s += (double)v.x;
      q += (double)v.y;
    }
    wsum[wid][lane][0] = s;
    wsum[wid][lane][1] = q;
    __syncthreads();
    if (tid < kGroups) {
      s = 0.0; q = 0.0;
      for (int w = 0; w < nwarps; ++w) { s += wsum[w][tid][0]; q += wsum[w][tid][1]; }
      const double cnt = (double)rows_per_unit * cg;
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      mean_s[tid] = (float)mean;
      rstd_s[tid] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }
  __syncthreads();
  if (tid >= nwork) return;
  const int vec = tid % vpr;
  const int rpar = nwork / vpr;
  const int rsub = tid / vpr;
  const int c = vec << 3;
  const bf16* base;
  int ld;
  if (c < C0) { base = src0 + c; ld = C0; } else { base = src1 + (c - C0); ld = C1; }
  // per-channel scale/shift (fold mean/rstd into gamma/beta)
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cg;
    const float ga = __ldg(gamma + c + j), be = __ldg(beta + c + j);
    sc[j] = rstd_s[g] * ga;
    sh[j] = be - mean_s[g] * sc[j];
  }
  const int rows_per_blk = (rows_per_unit + blocks_per_unit - 1) / blocks_per_unit;
  const int r_begin = blk * rows_per_blk;
  int r_end = r_begin + rows_per_blk;
  if (r_end > rows_per_unit) r_end = rows_per_unit;
  const size_t row0 = (size_t)unit * rows_per_unit;
#pragma unroll 8
  for (int r = r_begin + rsub; r < r_end; r += rpar) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + (row0 + r) * ld));
    float v[8];
    float2 f;
    f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
    f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
    f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
    f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = v[j] * sc[j] + sh[j];
      v[j] = silu ? silu_f(t) : t;
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + (row0 + r) * C + c) = o;
  }
}

// ------------------------------------------------------------------------------------------
// LayerNorm: a row is owned by LPR lanes (8/16/32) holding VPL 16-byte vectors each, so that a
// 320-channel row (40 vectors) keeps every lane busy; two-pass statistics in registers.
// ------------------------------------------------------------------------------------------
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, long long ldx, int M, int C,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, const float* __restrict__ rowbias, int ld_rowbias,
                                 int rb_div, int rb_mod, bf16* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gtid / LPR;
  const int sl = threadIdx.x % LPR;  // sub-lane inside the row group
  const bool active = row < M;
  const int vpr = C >> 3;
  const int rowc = active ? row : 0;
  const bf16* xr = x + (size_t)rowc * ldx;
  const float* rb = rowbias ? rowbias + (size_t)((rowc / rb_div) % rb_mod) * ld_rowbias : nullptr;
  float v[VPL][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = sl + LPR * i;
    if (vec < vpr) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(xr + vec * 8));
      float2 f;
      f = unpack_bf16x2(u.x); v[i][0] = f.x; v[i][1] = f.y;
      f = unpack_bf16x2(u.y); v[i][2] = f.x; v[i][3] = f.y;
      f = unpack_bf16x2(u.z); v[i][4] = f.x; v[i][5] = f.y;
      f = unpack_bf16x2(u.w); v[i][6] = f.x; v[i][7] = f.y;
      if (rb) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(rb + vec * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(rb + vec * 8 + 4));
        v[i][0] += a.x; v[i][1] += a.y; v[i][2] += a.z; v[i][3] += a.w;
        v[i][4] += b.x; v[i][5] += b.y; v[i][6] += b.z; v[i][7] += b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = sl + LPR * i;
    if (vec < vpr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + eps);
  if (!active) return;
  bf16* orow = out + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vec = sl + LPR * i;
    if (vec < vpr) {
      float o[8];
      const float nm = -mean * rstd;
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(v[i][j], rstd, nm);
      if (gamma) {  // affine folded into the consumer's weights when gamma == nullptr
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vec * 8));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vec * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vec * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vec * 8 + 4));
        o[0] = o[0] * g0.x + b0.x; o[1] = o[1] * g0.y + b0.y; o[2] = o[2] * g0.z + b0.z; o[3] = o[3] * g0.w + b0.w;
        o[4] = o[4] * g1.x + b1.x; o[5] = o[5] * g1.y + b1.y; o[6] = o[6] * g1.z + b1.z; o[7] = o[7] * g1.w + b1.w;
      }
      uint4 u;
      u.x = pack_bf16x2(o[0], o[1]); u.y = pack_bf16x2(o[2], o[3]);
      u.z = pack_bf16x2(o[4], o[5]); u.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(orow + vec * 8) = u;
    }
  }
}

}  // namespace ctrlv

using namespace ctrlv;

extern "C" int64_t ctrlv_groupnorm_workspace(int32_t n_units) {
  return (int64_t)n_units * kMaxSplit * kGroups * 2 * sizeof(float);
}

static int n_sm_current(int* out) {
  static int n_sm_of[64] = {0};  // per device ordinal
  int dev = 0;
  CTRLV_CUDA(cudaGetDevice(&dev));
  CTRLV_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  if (n_sm_of[dev] == 0) CTRLV_CUDA(cudaDeviceGetAttribute(&n_sm_of[dev], cudaDevAttrMultiProcessorCount, dev));
  *out = n_sm_of[dev];
  return CTRLV_OK;
}

// fixed_sums == nullptr: statistics pass into `workspace`, then apply; else apply only
static int groupnorm_impl(const void* src0, int32_t C0, const void* src1, int32_t C1,
                          int32_t n_units, int32_t rows_per_unit, const float* gamma,
                          const float* beta, float eps, int32_t silu, void* out,
                          void* workspace, const void* fixed_sums, int n_rep, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!src1) C1 = 0;
  const int C = C0 + C1;
  CTRLV_CHECK_ARG(src0 && out && (workspace || fixed_sums) && gamma && beta, "groupnorm: null pointer");
  CTRLV_CHECK_ARG(C0 % 8 == 0 && C1 % 8 == 0 && C % 64 == 0, "groupnorm: C0=%d C1=%d need %%8 and sum %%64", C0, C1);
  CTRLV_CHECK_ARG(n_units > 0 && rows_per_unit > 0, "groupnorm: empty input");
  const int vpr = C / 8;
  CTRLV_CHECK_ARG(vpr <= 1024, "groupnorm: C=%d too wide", C);
  int rpar = 512 / vpr;
  if (rpar < 1) rpar = 1;
  if (rpar > rows_per_unit) rpar = rows_per_unit;
  const int nwork = vpr * rpar;
  const int nthreads = (nwork + 31) / 32 * 32;
  // Both kernels do uniform work per block: size each grid to ONE full wave of resident blocks
  // (a 1.04-wave grid costs two waves).  Statistics splits and apply blocks are independent.
  int n_sm = 0;
  {
    const int rc = n_sm_current(&n_sm);
    if (rc) return rc;
  }
  const size_t st_smem = ((size_t)nwork * 8 + (size_t)vpr * 8) * sizeof(float);
  int occ_stats = 1, occ_apply = 1;
  CTRLV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_stats, gn_stats_kernel, nthreads, st_smem));
  CTRLV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_apply, gn_apply_kernel, nthreads, 0));
  if (occ_stats < 1) occ_stats = 1;
  if (occ_apply < 1) occ_apply = 1;
  const int max_by_rows = (rows_per_unit + rpar - 1) / rpar;
  auto per_unit = [&](int slots) {
    int n = slots / n_units;  // floor: never spill into a second wave
    if (n > max_by_rows) n = max_by_rows;
    if (n > kMaxSplit) n = kMaxSplit;
    if (n < 1) n = 1;
    return n;
  };
  int nsplit = per_unit(occ_stats * n_sm);
  const int rows_per_split = (rows_per_unit + nsplit - 1) / nsplit;
  nsplit = (rows_per_unit + rows_per_split - 1) / rows_per_split;
  int nblk = per_unit(occ_apply * n_sm);
  const int rows_per_blk = (rows_per_unit + nblk - 1) / nblk;
  nblk = (rows_per_unit + rows_per_blk - 1) / rows_per_blk;
  float* partial = reinterpret_cast<float*>(workspace);
  if (!fixed_sums)
    CTRLV_CUDA(launch_pdl(gn_stats_kernel, dim3(n_units * nsplit), dim3(nthreads), st_smem, stream,
                          reinterpret_cast<const bf16*>(src0), C0, reinterpret_cast<const bf16*>(src1), C1,
                          rows_per_unit, rows_per_split, nsplit, nwork, partial));
  CTRLV_CUDA(launch_pdl(gn_apply_kernel, dim3(n_units * nblk), dim3(nthreads), (size_t)0, stream,
                        reinterpret_cast<const bf16*>(src0), C0, reinterpret_cast<const bf16*>(src1), C1,
                        rows_per_unit, nblk, nsplit, nwork, (const float*)partial, gamma, beta, eps, silu,
                        reinterpret_cast<bf16*>(out), reinterpret_cast<const long long*>(fixed_sums), n_units, n_rep));
  return CTRLV_OK;
}

extern "C" int ctrlv_groupnorm(const void* src0, int32_t C0, const void* src1, int32_t C1,
                               int32_t n_units, int32_t rows_per_unit, const float* gamma,
                               const float* beta, float eps, int32_t silu, void* out,
                               void* workspace, void* stream_) {
  CTRLV_CHECK_ARG(workspace != nullptr, "groupnorm: null workspace");
  return groupnorm_impl(src0, C0, src1, C1, n_units, rows_per_unit, gamma, beta, eps, silu, out, workspace,
                        nullptr, 1, stream_);
}

extern "C" int ctrlv_groupnorm_apply(const void* src0, int32_t C0, const void* src1, int32_t C1,
                                     int32_t n_units, int32_t rows_per_unit, const float* gamma,
                                     const float* beta, float eps, int32_t silu, void* out,
                                     const void* sums, int32_t n_rep, void* stream_) {
  CTRLV_CHECK_ARG(sums != nullptr && (reinterpret_cast<uintptr_t>(sums) & 7) == 0, "groupnorm_apply: null or misaligned sums");
  CTRLV_CHECK_ARG(n_rep >= 1 && (n_rep & (n_rep - 1)) == 0, "groupnorm_apply: n_rep must be a power of two");
  return groupnorm_impl(src0, C0, src1, C1, n_units, rows_per_unit, gamma, beta, eps, silu, out, nullptr, sums,
                        n_rep, stream_);
}

extern "C" int ctrlv_axpby_gn(const void* x, const void* y, float a, float b, int64_t rows, int32_t C, void* out,
                              void* gn_sums, int32_t rows_per_unit, int32_t cg, int32_t c_off, int32_t n_rep,
                              void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(x && ((y == nullptr) == (out == nullptr)) && gn_sums && rows > 0 && C > 0 && C % 8 == 0 && C <= 8 * 1024,
                  "axpby_gn: bad arguments (y and out both or neither, C %% 8, C <= 8192)");
  CTRLV_CHECK_ARG(rows_per_unit > 0 && rows % rows_per_unit == 0 && cg >= 2 && cg % 2 == 0 && c_off >= 0 && c_off % 2 == 0,
                  "axpby_gn: rows %% rows_per_unit, even cg and c_off");
  CTRLV_CHECK_ARG((c_off + C + cg - 1) / cg <= kGroups, "axpby_gn: channels fall outside the 32 groups");
  CTRLV_CHECK_ARG(n_rep >= 1 && (n_rep & (n_rep - 1)) == 0, "axpby_gn: n_rep must be a power of two");
  const int n_units = (int)(rows / rows_per_unit);
  const int vpr = C / 8;
  int rpar = 512 / vpr;
  if (rpar < 1) rpar = 1;
  if (rpar > rows_per_unit) rpar = rows_per_unit;
  const int nwork = vpr * rpar;
  const int nthreads = (nwork + 31) / 32 * 32;
  int n_sm = 0;
  {
    const int rc = n_sm_current(&n_sm);
    if (rc) return rc;
  }
  const size_t smem = ((size_t)nwork * 8 + (size_t)vpr * 8) * sizeof(float);
  int occ = 1;
  CTRLV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, axpby_gn_kernel, nthreads, smem));
  if (occ < 1) occ = 1;
  int nsplit = occ * n_sm / n_units;  // one wave of resident blocks
  const int max_by_rows = (rows_per_unit + rpar - 1) / rpar;
  if (nsplit > max_by_rows) nsplit = max_by_rows;
  if (nsplit < 1) nsplit = 1;
  const int rows_per_split = (rows_per_unit + nsplit - 1) / nsplit;
  nsplit = (rows_per_unit + rows_per_split - 1) / rows_per_split;
  CTRLV_CUDA(launch_pdl(axpby_gn_kernel, dim3(n_units * nsplit), dim3(nthreads), smem, stream,
                        reinterpret_cast<const bf16*>(x), reinterpret_cast<const bf16*>(y), a, b, C, rows_per_unit,
                        rows_per_split, nsplit, nwork, reinterpret_cast<bf16*>(out),
                        reinterpret_cast<unsigned long long*>(gn_sums), cg, c_off, n_units, n_rep));
  return CTRLV_OK;
}

extern "C" int ctrlv_layernorm(const void* x, int64_t ldx, int32_t M, int32_t C, const float* gamma,
                               const float* beta, float eps, const float* rowbias,
                               int32_t ld_rowbias, int32_t rb_div, int32_t rb_mod, void* out,
                               void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  CTRLV_CHECK_ARG(x && out && ((gamma == nullptr) == (beta == nullptr)), "layernorm: null pointer");
  CTRLV_CHECK_ARG(C % 8 == 0 && C <= 8 * 32 * 8, "layernorm: C=%d unsupported", C);
  CTRLV_CHECK_ARG(ldx % 8 == 0, "layernorm: ldx %% 8");
  if (rowbias) CTRLV_CHECK_ARG(rb_div > 0 && rb_mod > 0 && ld_rowbias % 4 == 0, "layernorm: bad rowbias args");
  const int vpr = C / 8;
  // lanes per row / vectors per lane: keep all lanes busy for C = 320 (8x5), 640 (16x5), 1280 (32x5)
  int lpr = 32, vpl = 8;
  if (vpr <= 8) { lpr = 8; vpl = 1; }
  else if (vpr <= 16) { lpr = 8; vpl = 2; }
  else if (vpr <= 40) { lpr = 8; vpl = 5; }
  else if (vpr <= 80) { lpr = 16; vpl = 5; }
  else if (vpr <= 160) { lpr = 32; vpl = 5; }
  const int threads = 256;
  const long long total = (long long)M * lpr;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads);
#define CTRLV_LN(L, V)                                                                             \
  CTRLV_CUDA(launch_pdl(layernorm_kernel<L, V>, dim3(blocks), dim3(threads), (size_t)0, stream,        \
                        reinterpret_cast<const bf16*>(x), (long long)ldx, M, C, gamma, beta, eps, rowbias, \
                        ld_rowbias, rb_div, rb_mod, reinterpret_cast<bf16*>(out)))
  if (lpr == 8 && vpl == 1) CTRLV_LN(8, 1);
  else if (lpr == 8 && vpl == 2) CTRLV_LN(8, 2);
  else if (lpr == 8 && vpl == 5) CTRLV_LN(8, 5);
  else if (lpr == 16) CTRLV_LN(16, 5);
  else if (vpl == 5) CTRLV_LN(32, 5);
  else CTRLV_LN(32, 8);
#undef CTRLV_LN
  CTRLV_CUDA(cudaGetLastError());
  return CTRLV_OK;
}
