// Fused FeedForward of the transformer blocks (diffusers FeedForward = GEGLU proj -> Linear, reached from
// BasicTransformerBlock.ff / TemporalBasicTransformerBlock.ff_in / .ff) for narrow models (C <= 320:
// level 0 of the SVD UNet, where the two-kernel form is bound by the up-projection's epilogue and writes a
// [rows x 4C] intermediate that the down-projection reads back):
//
//   out = s_acc * (GEGLU(x W1^T + b1) W2^T + b2 + rowbias[ridx(m)]) + s_res1 * res1 + s_res2 * res2
//
// One persistent CTA per SM walks 128-row tiles.  Per tile the x tile [128 x C] stays resident in shared
// memory; the hidden dimension is processed in chunks of 64 units:
//   MMA1  S[128 x 128]   = x W1_j^T            (smem x smem, fp32 in TMEM; value/gate columns interleaved)
//   GEGLU S -> H[128 x 64] bf16                (two groups of 8 epilogue warps on alternate chunks:
//                                               TMEM -> registers -> TMEM, packed bf16x2)
//   MMA2  D[128 x C]    += H W2_j^T            (A operand straight from TMEM, like P in the attention kernel)
// so the 4C-wide intermediate never leaves the SM, the second contraction runs under the GEGLU math of the
// next chunk, and there is one tile set-up and one output epilogue per 128 rows instead of eleven.
// W1_j / W2_j stream through a TMA ring in the order the MMA warp consumes them:
//   W1(0), W1(1), then for j = 0..J-1: [W1(j+2)], W2(j)   (MMA1 runs two chunks ahead of MMA2).
// TMEM columns: D [0, C), S [C, C+128), H double-buffered [C+128, C+192)  (C = 320: exactly 512).
//
// LayerNorm mode (ctrlv_feedforward_ln): x holds the rows BEFORE the nn.LayerNorm in front of the FeedForward.  The
// resident x tile has whole rows, so the 16 epilogue warps normalise it in place (4 threads per row, three passes over
// the swizzled tile, the fp32 arithmetic of layernorm_kernel) and release the MMA warp through `x_ready`; the tile of
// the next unit is normalised between this unit's last GEGLU chunk and its output epilogue.  In this mode every CTA
// of a pair receives its x tile on its own barrier (`x_land`) instead of the leader's.
//
// Projection mode (ctrlv_linear_ln: LayerNorm + Linear, e.g. norm1 + to_q | to_k | to_v): the same launch without
// GEGLU, H, MMA2 and D — W1 is the projection's [N][C] matrix, S is double-buffered in the TMEM columns D does not
// need, and each 128-column chunk leaves the launch from S as bf16 rows (+ bias).
//
// CG = 2 (the production form): a CTA pair (cluster of 2, tcgen05 cta_group::2) works on two adjacent row
// tiles at once; each CTA loads HALF of every weight tile and the leader's M = 256 MMAs read both halves.
// One weight chunk is 120 KB, exactly the shared memory left beside the resident x tile, so a single CTA
// exposes one TMA round trip per chunk (measured: 4600 cycles per chunk against 1920 of MMA work); the pair
// halves the bytes per CTA, which makes the ring two chunks deep, and halves the L2 -> SM weight traffic.
#include "common.cuh"
#include "../../include/ctrlv_b200.h"

namespace ctrlv {

constexpr int kFfEpiWarps = 16;  // four per TMEM lane quarter: the GEGLU stage is a latency chain per warp
                                 // (TMEM load -> bias -> polynomial -> tanh -> pack -> TMEM store), so its pace is
                                 // set by how short each warp's share of a chunk is, not by issue slots
constexpr int kFfSub = kFfEpiWarps / 4;
constexpr int kFfThreads = (4 + kFfEpiWarps) * 32;  // warp 0: TMA producer, 1: MMA issuer (+ TMEM alloc), 2-3 idle
constexpr int kFfMaxStages = 12;
constexpr int kFfXBlock = 128 * 64 * 2;  // one 64-channel k-block of the x tile / of a W1 chunk

struct FfParams {
  CUtensorMap tmX, tmW1, tmW2;
  int M, C, H;        // H = 4C hidden units; W1 has 2H rows (value, gate interleaved)
  int KB1;            // C / 64 k-blocks of MMA1
  int J;              // H / 64 hidden chunks
  int n2, ntile2;     // MMA2 n-tile width (<= 256) and count: n2 * ntile2 == C
  int cg;             // CTAs per tile group (1 or 2)
  int stages, stage_bytes, tiles;
  int off_ring, off_b1, off_b2;  // byte offsets inside dynamic shared memory
  const float* b1;
  // ln != 0: x holds the rows BEFORE the nn.LayerNorm(C, ln_eps) that feeds this FeedForward (its affine part is
  // folded into W1 / b1); every x tile is normalised in shared memory before the first contraction reads it.
  // ln_rb: optional fp32 row-bias added before the statistics, row m takes ln_rb[(m / ln_rb_div) % ln_rb_mod]
  // proj != 0 (ctrlv_linear_ln): no GEGLU, no second contraction — out[m][n] = LN(x)[m] . W1[n] + b1[n] for N
  // outputs, written chunk by chunk from S (W1 = the projection's [N][C] weights, J = ceil(N / 128) chunks); the x
  // tile is double-buffered (nxb = 2) so that tile i + 1 is loaded and normalised under the chunks of tile i
  int proj, N, nxb;
  int ln;
  float ln_eps;
  const float* ln_rb;
  int ln_rb_ld, ln_rb_div, ln_rb_mod;
  ctrlv_epilogue ep;  // output epilogue: bias (= b2), rowbias, s_acc, res1, res2, out (bf16)
};

__device__ __forceinline__ int ff_rowbias_index(const ctrlv_epilogue& ep, int m) {
  const int a = m / ep.rb_div;
  if (ep.rb_mode == 1) return a;
  if (ep.rb_mode == 2) return a % ep.rb_mod;
  return (a * ep.rb_mod + m % ep.rb_mod + ep.rb_off) % ep.rb_B;
}

__device__ __forceinline__ uint4 lds_u4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_u4(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {  // (the dynamic-smem pointers are generic: be explicit)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
// spin on the barrier phase without the hardware-suspend hint of mbar_wait: the handoffs of this kernel
// (S full -> TMEM load -> S free -> MMA1, H full -> MMA2) sit on the critical path several times per chunk
__device__ __forceinline__ void ff_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void ff_add_bf16x16(float* v, const uint4& lo, const uint4& hi, float s) {
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float2 f = unpack_bf16x2(w[k]);
    v[2 * k] = fmaf(s, f.x, v[2 * k]);
    v[2 * k + 1] = fmaf(s, f.y, v[2 * k + 1]);
  }
}

template <int CG>
__global__ void __launch_bounds__(kFfThreads, 1) ff_kernel(const __grid_constant__ FfParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t x_full[2], x_empty[2], d_full, d_free;  // x barriers: one per x buffer (p.nxb)
  __shared__ __align__(8) uint64_t x_land[2], x_ready[2];  // LayerNorm mode: x tile landed (per CTA) / normalised (on the leader)
  __shared__ __align__(8) uint64_t s_full[2], s_free[2], h_full[2], h_free[2];  // indexed by chunk parity
  __shared__ __align__(8) uint64_t full_bar[kFfMaxStages], empty_bar[kFfMaxStages];
  __shared__ uint32_t tmem_base_smem;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sX = smem;
  uint8_t* ring = smem + p.off_ring;
  float* sb1 = reinterpret_cast<float*>(smem + p.off_b1);
  float* sb2 = reinterpret_cast<float*>(smem + p.off_b2);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
      mbar_init(&x_land[i], 1);
      mbar_init(&x_ready[i], kFfEpiWarps * CG);
    }
    mbar_init(&d_full, 1);
    mbar_init(&d_free, kFfEpiWarps * CG);  // CG = 2: both CTAs' epilogue warps report to the leader
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], kFfEpiWarps / 2 * CG);
      mbar_init(&h_full[i], kFfEpiWarps / 2 * CG);
      mbar_init(&h_free[i], 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 1) tmem_alloc(&tmem_base_smem, 512);
    else tmem_alloc_cg2(&tmem_base_smem, 512);
  }
  // biases are weights (not written by the predecessor kernel): staged once per CTA
  {
    const int nb1 = p.proj ? p.N : 2 * p.H;  // (projection: its bias, zero past N)
    for (int i = threadIdx.x; i < p.J * 128; i += kFfThreads) sb1[i] = (p.b1 != nullptr && i < nb1) ? __ldg(p.b1 + i) : 0.f;
    if (!p.proj)
      for (int i = threadIdx.x; i < p.C; i += kFfThreads) sb2[i] = p.ep.bias ? __ldg(p.ep.bias + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // peer barriers initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t crank = (CG == 2) ? cluster_ctarank() : 0u;
  pdl_wait();

  // work unit: CG adjacent row tiles; units are dealt round-robin to the CTA groups, CTA r of a group owns
  // tile CG * unit + r (past the last tile: TMA zero fill, nothing stored)
  const int nunits = (p.tiles + CG - 1) / CG, ngroups = (int)gridDim.x / CG, group = (int)blockIdx.x / CG;
  const int nloc = group < nunits ? (nunits - group + ngroups - 1) / ngroups : 0;
  auto tile_row0 = [&](int it) { return ((group + it * ngroups) * CG + (int)crank) * 128; };
  const uint32_t tD = tmem_base, tS = tmem_base + (uint32_t)p.C, tH = tmem_base + (uint32_t)p.C + 128u;
  // x buffer / barrier phase of this CTA's it-th unit
  auto xbuf = [&](int it) { return p.nxb == 2 ? (it & 1) : 0; };
  auto xpar = [&](int it) { return (uint32_t)((p.nxb == 2 ? (it >> 1) : it) & 1); };
  const int xbytes = p.KB1 * kFfXBlock;

  // register reallocation between the warpgroups (each executes ONE setmaxnreg at the head of its role branch).
  // Budget: 640 threads x 96 registers at launch; warps 0-3 release 128 x (96 - 40) = 7168, the four epilogue
  // warpgroups take 4 x 128 x (104 - 96) = 4096 (asking for more than was released never returns).
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    // One ring slot holds a whole operand group — all KB1 k-blocks of a W1 chunk, or all n-tiles of a W2 chunk —
    // so the MMA warp pays one barrier wait / elect / commit round per group instead of one per 16-32 KB block:
    // with a round per block its own instruction stream (~75 instructions, ~600 cycles per round) outlasted the
    // 256 cycles of tensor work a block carries, in every organisation of the other warps.
    // `rows` weight rows per block in total; each CTA of a pair loads rows / CG of them, starting at its share.
    // CG = 2: both CTAs credit the LEADER's full barrier, which expects the pair's bytes.
    auto load_group = [&](const CUtensorMap* tm, int nblk, int rows, bool w1, int j) {
      mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* dst = ring + (size_t)stage * p.stage_bytes;
        const int blk_bytes = (rows / CG) * 128;
        const uint32_t lead_bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
        if (CG == 1 || crank == 0) mbar_expect_tx(&full_bar[stage], (uint32_t)(nblk * rows * 128));
        for (int b = 0; b < nblk; ++b) {
          const int c0 = w1 ? b * 64 : j * 64;                     // K coordinate (W1: channel block, W2: hidden chunk)
          const int r0 = (w1 ? j * 128 : b * rows) + (int)crank * (rows / CG);
          if (CG == 1) tma_load_2d(dst + (size_t)b * blk_bytes, tm, &full_bar[stage], c0, r0);
          else tma_load_2d_cg2(dst + (size_t)b * blk_bytes, tm, lead_bar, c0, r0);
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    };
    auto w1 = [&](int j) { load_group(&p.tmW1, p.KB1, 128, true, j); };
    auto w2 = [&](int j) { load_group(&p.tmW2, p.ntile2, p.n2, false, j); };
    auto load_x = [&](int it) {  // the x tile of this CTA's it-th unit, once MMA1 of the previous unit is done with x
      const int m0 = tile_row0(it);
      const int xb = xbuf(it);
      uint8_t* dstx = sX + (size_t)xb * xbytes;
      mbar_wait_relaxed(&x_empty[xb], xpar(it) ^ 1u);
      if (elect_one()) {
        if (p.ln) {  // each CTA's tile reports to its OWN barrier: its epilogue warps normalise it before the MMAs
          mbar_expect_tx(&x_land[xb], (uint32_t)xbytes);
          for (int kb = 0; kb < p.KB1; ++kb) tma_load_2d(dstx + (size_t)kb * kFfXBlock, &p.tmX, &x_land[xb], kb * 64, m0);
        } else if (CG == 1) {
          mbar_expect_tx(&x_full[xb], (uint32_t)xbytes);
          for (int kb = 0; kb < p.KB1; ++kb) tma_load_2d(dstx + (size_t)kb * kFfXBlock, &p.tmX, &x_full[xb], kb * 64, m0);
        } else {
          const uint32_t lead_bar = smem_u32(&x_full[xb]) & 0xFEFFFFFFu;
          if (crank == 0) mbar_expect_tx(&x_full[xb], (uint32_t)(2 * xbytes));
          for (int kb = 0; kb < p.KB1; ++kb) tma_load_2d_cg2(dstx + (size_t)kb * kFfXBlock, &p.tmX, lead_bar, kb * 64, m0);
        }
      }
      __syncwarp();
    };
    // x first, then the weight chunks in the MMA warp's order (x BEFORE the next unit's first chunks: those sit
    // in the ring until MMA1 can read x, a ring shorter than two chunks would otherwise never drain)
    if (nloc > 0) {
      load_x(0);
      w1(0);
      if (p.J > 1) w1(1);
    }
    for (int it = 0; it < nloc; ++it) {
      if (p.nxb == 2 && it + 1 < nloc) load_x(it + 1);  // second x buffer: the next unit's rows land a whole unit early
      for (int j = 0; j < p.J; ++j) {
        if (j + 2 < p.J) w1(j + 2);
        if (!p.proj) w2(j);
      }
      if (it + 1 < nloc) {  // the next unit's x and first chunks stream in under this unit's tail
        if (p.nxb != 2) load_x(it + 1);
        w1(0);
        if (p.J > 1) w1(1);
      }
    }
  } else if (warp == 1 && crank == 0) {
    // ================================ MMA issuer (warp-uniform, elected lane issues) ============
    const uint32_t idesc1 = make_idesc(128 * CG, 128, 0, 0);
    const uint32_t idesc2 = make_idesc(128 * CG, p.n2, 0, 0);  // A (= H) from TMEM
    auto commit = [&](uint64_t* bar) {  // CG = 2: the arrive lands on the barrier at this offset in BOTH CTAs
      if (CG == 1) umma_commit(bar);
      else umma_commit_cg2(bar);
    };
    const uint32_t aR = smem_u32(ring);
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < nloc; ++it) {
      auto mma1 = [&](int j) {  // S = x W1_j^T
        const int g = it * p.J + j;
        // projection mode: D is unused, so S is double-buffered (columns [0, 128) / [128, 256)) and chunk g only waits
        // for chunk g - 2 to have left ITS buffer: the up-projections run back to back under the stores
        const uint32_t tSg = p.proj ? tmem_base + (uint32_t)((g & 1) * 128) : tS;
        if (p.proj) {
          if (g >= 2) {
            ff_wait(&s_free[g & 1], (uint32_t)(((g >> 1) - 1) & 1));
            tc_fence_after();
          }
        } else if (g > 0) {  // S of the previous chunk sits in its epilogue group's registers
          ff_wait(&s_free[(g - 1) & 1], (uint32_t)(((g - 1) >> 1) & 1));
          tc_fence_after();
        }
        ff_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da0 = make_sdesc(smem_u32(sX) + (uint32_t)(xbuf(it) * xbytes), 16, 1024);
          const uint64_t db0 = make_sdesc(aR + (uint32_t)(stage * p.stage_bytes), 16, 1024);
          for (int kb = 0; kb < p.KB1; ++kb) {
            // k-block kb of x / of this CTA's share of the W1 chunk; +2 in the >>4 address field = 32 bytes = K 16
            const uint64_t da = da0 + (uint64_t)((kb * kFfXBlock) >> 4);
            const uint64_t db = db0 + (uint64_t)((kb * (128 / CG) * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CG == 1) umma_ss(tSg, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (uint32_t)((kb | k) != 0));
              else umma_ss_cg2(tSg, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc1, (uint32_t)((kb | k) != 0));
            }
          }
          commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (elect_one()) {
          commit(&s_full[g & 1]);
          if (j == p.J - 1) commit(&x_empty[xbuf(it)]);
        }
        __syncwarp();
      };
      auto mma2 = [&](int j) {  // D (+)= H_j W2_j^T
        const int g = it * p.J + j;
        const int buf = g & 1;
        ff_wait(&h_full[buf], (uint32_t)((g >> 1) & 1));
        tc_fence_after();
        if (j == 0) {  // the output epilogue of the previous tile has read D
          ff_wait(&d_free, (uint32_t)((it & 1) ^ 1));
          tc_fence_after();
        }
        ff_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t db0 = make_sdesc(aR + (uint32_t)(stage * p.stage_bytes), 16, 1024);
          for (int t = 0; t < p.ntile2; ++t) {
            const uint64_t db = db0 + (uint64_t)((t * (p.n2 / CG) * 128) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (CG == 1)
                umma_ts(tD + (uint32_t)(t * p.n2), tH + (uint32_t)(buf * 32 + k * 8), db + (uint64_t)(2 * k), idesc2,
                        (uint32_t)((j | k) != 0));
              else
                umma_ts_cg2(tD + (uint32_t)(t * p.n2), tH + (uint32_t)(buf * 32 + k * 8), db + (uint64_t)(2 * k), idesc2,
                            (uint32_t)((j | k) != 0));
            }
          }
          commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (elect_one()) {
          commit(&h_free[buf]);
          if (j == p.J - 1) commit(&d_full);
        }
        __syncwarp();
      };
      ff_wait(p.ln ? &x_ready[xbuf(it)] : &x_full[xbuf(it)], xpar(it));
      tc_fence_after();
      // MMA1 runs TWO chunks ahead of MMA2: S(j + 2) only waits for the epilogue to have pulled S(j + 1) out of
      // TMEM, so the S hand-off loop never queues behind a second contraction that is still waiting for its
      // GEGLU output, and MMA2(j) fills the tensor pipe during the next hand-off instead
      mma1(0);
      if (p.J > 1) mma1(1);
      for (int j = 0; j < p.J; ++j) {
        if (j + 2 < p.J) mma1(j + 2);
        if (!p.proj) mma2(j);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ================================ epilogue: GEGLU per chunk, output per tile ================
    const int q = warp & 3;           // TMEM lane quarter
    const int sub = (warp - 4) >> 2;  // output chunks c = sub, sub + 4, ...
    const int grp = sub >> 1;         // the two groups of 8 warps take alternate hidden chunks, so that chunk
                                      // g + 1 is pulled out of S (freeing it for MMA1 of g + 2) while chunk g is
                                      // still in its GEGLU math
    const int half = sub & 1;         // which 64 of the chunk's 128 columns
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t sb1_addr = smem_u32(sb1), sb2_addr = smem_u32(sb2);
    const ctrlv_epilogue& ep = p.ep;
    auto arrive = [&](uint64_t* bar) {  // on the leader CTA's barrier
      if (CG == 1) mbar_arrive(bar);
      else mbar_arrive_cluster(mapa_u32(smem_u32(bar), 0));
    };
    // LayerNorm of the x tile in place (p.ln): 4 threads per row (8 rows per warp), three passes over the row in
    // shared memory — sum, centred sum of squares, normalise — the arithmetic of layernorm_kernel (fp32 two-pass
    // statistics, one rounding to bf16), so the tile the MMAs read equals what the separate launch would have
    // written.  16-byte piece pc of row r sits at k-block pc / 8, byte ((pc % 8) ^ (r % 8)) * 16 of the 128-byte row
    // (TMA 128B swizzle).  Generic-proxy writes are fenced for the async proxy before the MMA warp is released.
    auto normalise_x = [&](int it) {
      ff_wait(&x_land[xbuf(it)], xpar(it));
      const int row = (warp - 4) * 8 + (lane >> 2), jq = lane & 3;
      const int npt = p.C >> 5;  // pieces per thread: pc = 4 * i + jq
      const int m = tile_row0(it) + row;
      const uint32_t xrow = smem_u32(sX) + (uint32_t)(xbuf(it) * xbytes) + (uint32_t)(row * 128);
      const float* rbp = (p.ln_rb != nullptr && m < p.M)
                             ? p.ln_rb + (size_t)((m / p.ln_rb_div) % p.ln_rb_mod) * p.ln_rb_ld : nullptr;
      auto load8 = [&](int pc, float* v) {
        const uint4 u = lds_u4(xrow + (uint32_t)((pc >> 3) * kFfXBlock) + (uint32_t)((((pc & 7) ^ (row & 7))) << 4));
        float2 f;
        f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
        f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
        f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
        f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
        if (rbp) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(rbp + pc * 8));
          const float4 b = __ldg(reinterpret_cast<const float4*>(rbp + pc * 8 + 4));
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
      };
      float sum = 0.f;
      for (int i = 0; i < npt; ++i) {
        float v[8];
        load8(4 * i + jq, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum += v[k];
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      const float mean = sum / (float)p.C;
      float sq = 0.f;
      for (int i = 0; i < npt; ++i) {
        float v[8];
        load8(4 * i + jq, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = v[k] - mean;
          sq += d * d;
        }
      }
      sq += __shfl_xor_sync(0xffffffffu, sq, 1);
      sq += __shfl_xor_sync(0xffffffffu, sq, 2);
      const float rstd = rsqrtf(sq / (float)p.C + p.ln_eps);
      const float nm = -mean * rstd;
      for (int i = 0; i < npt; ++i) {
        const int pc = 4 * i + jq;
        float v[8];
        load8(pc, v);
        uint4 o;
        o.x = pack_bf16x2(fmaf(v[0], rstd, nm), fmaf(v[1], rstd, nm));
        o.y = pack_bf16x2(fmaf(v[2], rstd, nm), fmaf(v[3], rstd, nm));
        o.z = pack_bf16x2(fmaf(v[4], rstd, nm), fmaf(v[5], rstd, nm));
        o.w = pack_bf16x2(fmaf(v[6], rstd, nm), fmaf(v[7], rstd, nm));
        sts_u4(xrow + (uint32_t)((pc >> 3) * kFfXBlock) + (uint32_t)((((pc & 7) ^ (row & 7))) << 4), o);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) arrive(&x_ready[xbuf(it)]);
    };
    if (p.ln && nloc > 0) normalise_x(0);
    if (p.proj) {
      // ================= projection mode: S chunk -> + bias -> bf16 rows of out (no GEGLU, no second contraction)
      const int jn = p.J >> 1;  // the next tile (second x buffer) is normalised half-way through this one
      for (int it = 0; it < nloc; ++it) {
        const int m = tile_row0(it) + q * 32 + lane;
        // (loop-invariant tests as predicates: a per-chunk reload of a spilled bound queued behind this loop's
        // stores in the L1 pipeline and cost a third of the kernel's stall samples)
        const bool norm_next = p.ln && it + 1 < nloc;
        const bool norm_mid = norm_next && p.nxb == 2;
        bf16* orow = (m < p.M) ? reinterpret_cast<bf16*>(ep.out) + (size_t)m * ep.ld_out + half * 64 : nullptr;
        for (int j = 0; j < p.J; ++j) {
          if (norm_mid && j == jn) normalise_x(it + 1);  // (both groups, before their own chunk test)
          const int g = it * p.J + j;
          if ((g & 1) != grp) continue;
          ff_wait(&s_full[grp], (uint32_t)((g >> 1) & 1));
          tc_fence_after();
          uint32_t ra[32], rb[32];
          const uint32_t tSg = tmem_base + (uint32_t)(grp * 128) + lane_off;  // (g & 1 == grp: this group's S buffer)
          tmem_ld32(tSg + (uint32_t)(half * 64), ra);
          tmem_ld32(tSg + (uint32_t)(half * 64 + 32), rb);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive(&s_free[grp]);
          const int col0 = j * 128 + half * 64;
          if (orow != nullptr && col0 < p.N) {  // (N % 64 == 0: a 64-column half is stored whole or not at all)
            const uint32_t bb = sb1_addr + (uint32_t)(col0 * 4);
            bf16* op = orow + j * 128;
#pragma unroll
            for (int h = 0; h < 4; ++h) {  // 16 columns = one 32-byte sector per store
              const uint32_t* r = (h < 2 ? ra : rb) + (h & 1) * 16;
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 16; i += 4) {
                const float4 bv = lds_f4(bb + (uint32_t)((h * 16 + i) * 4));
                pk[i >> 1] = pack_bf16x2(__uint_as_float(r[i]) + bv.x, __uint_as_float(r[i + 1]) + bv.y);
                pk[(i >> 1) + 1] = pack_bf16x2(__uint_as_float(r[i + 2]) + bv.z, __uint_as_float(r[i + 3]) + bv.w);
              }
              st_global_v8(op + h * 16, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        // one x buffer: it is refilled only after this tile's last up-projection, normalise it now
        if (norm_next && !norm_mid) normalise_x(it + 1);
      }
    } else
    for (int it = 0; it < nloc; ++it) {
      const int m = tile_row0(it) + q * 32 + lane;
      const bool valid = m < p.M;
      for (int j = 0; j < p.J; ++j) {
        const int g = it * p.J + j;
        if ((g & 1) != grp) continue;
        const uint32_t ph = (uint32_t)((g >> 1) & 1);
        ff_wait(&s_full[grp], ph);
        tc_fence_after();
        uint32_t ra[32], rb[32];
        tmem_ld32(tS + lane_off + (uint32_t)(half * 64), ra);
        tmem_ld32(tS + lane_off + (uint32_t)(half * 64 + 32), rb);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(&s_free[grp]);
        // columns j*128 + half*64 + [0, 64): 32 (value, gate) pairs -> 32 hidden units
        const uint32_t bb = sb1_addr + (uint32_t)((j * 128 + half * 64) * 4);
        uint32_t pk[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t* r = h ? rb : ra;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = lds_f4(bb + (uint32_t)((h * 32 + i) * 4));
            float o0, o1;
            geglu2_f(__uint_as_float(r[i]) + b.x, __uint_as_float(r[i + 1]) + b.y, __uint_as_float(r[i + 2]) + b.z,
                     __uint_as_float(r[i + 3]) + b.w, o0, o1);
            pk[h * 8 + (i >> 2)] = pack_bf16x2(o0, o1);
          }
        }
        ff_wait(&h_free[grp], ph ^ 1);  // MMA2 of chunk g - 2 has consumed this buffer
        tc_fence_after();
        tmem_st16(tH + lane_off + (uint32_t)(grp * 32 + half * 16), pk);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive(&h_full[grp]);
      }
      // the next tile's x lands under this tile's last contractions: normalise it first, so that the next tile's
      // up-projections run under the output epilogue below
      if (p.ln && it + 1 < nloc) normalise_x(it + 1);  // (one x buffer in this mode: it has just been refilled)
      // ---- output epilogue of the tile: D + b2 + rowbias, scale, residuals, bf16 store (row-owner 32-byte accesses)
      const float* rbp = nullptr;
      if (ep.rb_mode != 0 && valid) rbp = ep.rowbias + (size_t)ff_rowbias_index(ep, m) * ep.ld_rowbias;
      ff_wait(&d_full, (uint32_t)(it & 1));
      tc_fence_after();
      for (int c = sub; c < p.C / 32; c += kFfSub) {
        uint32_t raw[32];
        tmem_ld32(tD + lane_off + (uint32_t)(c * 32), raw);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = lds_f4(sb2_addr + (uint32_t)((c * 32 + i) * 4));
          v[i] = __uint_as_float(raw[i]) + b.x; v[i + 1] = __uint_as_float(raw[i + 1]) + b.y;
          v[i + 2] = __uint_as_float(raw[i + 2]) + b.z; v[i + 3] = __uint_as_float(raw[i + 3]) + b.w;
        }
        if (rbp) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(rbp + c * 32 + i));
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
          }
        }
        if (ep.s_acc != 1.0f) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] *= ep.s_acc;
        }
        if (valid) {
          if (ep.res1) {
            const bf16* rp = reinterpret_cast<const bf16*>(ep.res1) + (size_t)m * ep.ld_res1 + c * 32;
            uint4 a0, a1, a2, a3;
            ld_global_nc_v8(rp, a0, a1);
            ld_global_nc_v8(rp + 16, a2, a3);
            ff_add_bf16x16(v, a0, a1, ep.s_res1);
            ff_add_bf16x16(v + 16, a2, a3, ep.s_res1);
          }
          if (ep.res2) {
            const bf16* rp = reinterpret_cast<const bf16*>(ep.res2) + (size_t)m * ep.ld_res2 + c * 32;
            uint4 a0, a1, a2, a3;
            ld_global_nc_v8(rp, a0, a1);
            ld_global_nc_v8(rp + 16, a2, a3);
            ff_add_bf16x16(v, a0, a1, ep.s_res2);
            ff_add_bf16x16(v + 16, a2, a3, ep.s_res2);
          }
          bf16* op = reinterpret_cast<bf16*>(ep.out) + (size_t)m * ep.ld_out + c * 32;
          st_global_v8(op, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                       pack_bf16x2(v[6], v[7]), pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]),
                       pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
          st_global_v8(op + 16, pack_bf16x2(v[16], v[17]), pack_bf16x2(v[18], v[19]), pack_bf16x2(v[20], v[21]),
                       pack_bf16x2(v[22], v[23]), pack_bf16x2(v[24], v[25]), pack_bf16x2(v[26], v[27]),
                       pack_bf16x2(v[28], v[29]), pack_bf16x2(v[30], v[31]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive(&d_free);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer may still read our smem / arrive on our barriers
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    if (CG == 1) tmem_dealloc(tmem_base, 512);
    else tmem_dealloc_cg2(tmem_base, 512);
  }
}

struct FfDev { int sms; int max_smem; };
static FfDev g_ff[64];
static int g_ff_force_cg = 0;

}  // namespace ctrlv

using namespace ctrlv;

// proj_N > 0: projection mode (ctrlv_linear_ln): W1 = [proj_N][C] weights, b1 = its bias (or NULL), W2 unused
static int feedforward_impl(const void* x, int64_t ldx, int32_t M, int32_t C, int ln, float ln_eps, const float* ln_rb,
                            int32_t ln_rb_ld, int32_t ln_rb_div, int32_t ln_rb_mod, const void* W1, const float* b1,
                            const void* W2, const ctrlv_epilogue* ep, void* stream_, int32_t proj_N = 0) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const bool proj = proj_N > 0;
  if (proj) W2 = W1;  // (alignment checks below)
  CTRLV_CHECK_ARG(x && W1 && (b1 || proj) && W2 && ep, "feedforward: null argument");
  CTRLV_CHECK_ARG(M > 0 && C > 0 && C % 64 == 0 && C <= 320, "feedforward: C=%d must be a multiple of 64, <= 320 (TMEM: C + 192 columns)", C);
  CTRLV_CHECK_ARG(ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W1) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(W2) & 15) == 0, "feedforward: operands must be 16-byte aligned");
  CTRLV_CHECK_ARG(ep->out != nullptr && ep->out_f32 == nullptr && ep->geglu == 0 && ep->n_store == 0 && ep->gn_sums == nullptr,
                  "feedforward: the output epilogue takes bias, rowbias, s_acc, res1, res2 and a bf16 out");
  CTRLV_CHECK_ARG(ep->ld_out % 16 == 0 && (reinterpret_cast<uintptr_t>(ep->out) & 31) == 0, "feedforward: out must be 32-byte aligned (ld %% 16)");
  if (ep->res1) CTRLV_CHECK_ARG(ep->ld_res1 % 16 == 0 && (reinterpret_cast<uintptr_t>(ep->res1) & 31) == 0, "feedforward: res1 must be 32-byte aligned");
  if (ep->res2) CTRLV_CHECK_ARG(ep->ld_res2 % 16 == 0 && (reinterpret_cast<uintptr_t>(ep->res2) & 31) == 0, "feedforward: res2 must be 32-byte aligned");
  if (ep->rb_mode != 0) {
    CTRLV_CHECK_ARG(ep->rowbias != nullptr && ep->rb_div > 0 && ep->ld_rowbias % 4 == 0, "feedforward: rowbias mode without table");
    if (ep->rb_mode == 2 || ep->rb_mode == 3) CTRLV_CHECK_ARG(ep->rb_mod > 0 && (ep->rb_mode == 2 || ep->rb_B > 0), "feedforward: bad rowbias modulus");
  }
  int dev = 0;
  CTRLV_CUDA(cudaGetDevice(&dev));
  CTRLV_CHECK_ARG(dev >= 0 && dev < 64, "device ordinal %d out of range", dev);
  FfDev& dp = g_ff[dev];
  if (dp.sms == 0) {
    int sms = 0, smem = 0;
    CTRLV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CTRLV_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    cudaFuncAttributes fa;
    CTRLV_CUDA(cudaFuncGetAttributes(&fa, ff_kernel<2>));
    smem -= (int)fa.sharedSizeBytes;
    CTRLV_CUDA(cudaFuncSetAttribute(ff_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CTRLV_CUDA(cudaFuncSetAttribute(ff_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dp.max_smem = smem;
    dp.sms = sms;
  }
  FfParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.C = C; p.H = 4 * C;
  p.KB1 = C / 64;
  p.J = proj ? (proj_N + 127) / 128 : p.H / 64;
  p.proj = proj ? 1 : 0; p.N = proj_N; p.nxb = 1;
  p.n2 = C > 256 ? C / 2 : C;
  p.ntile2 = C / p.n2;
  CTRLV_CHECK_ARG(p.n2 % 16 == 0, "feedforward: MMA2 n-tile %d", p.n2);
  p.tiles = (M + 127) / 128;
  // CTA pairs whenever there are two tiles and the weight halves keep whole 8-row swizzle groups
  p.cg = (g_ff_force_cg ? g_ff_force_cg : ((p.tiles >= 2 && p.n2 % 16 == 0) ? 2 : 1));
  CTRLV_CHECK_ARG(p.cg == 1 || p.n2 % 16 == 0, "feedforward: cta_group 2 needs an MMA2 n-tile that is a multiple of 16");
  {  // one ring slot = one operand group: a whole W1 chunk (KB1 k-blocks) or a whole W2 chunk (ntile2 n-tiles)
    const int g1 = p.KB1 * (128 / p.cg) * 128, g2 = proj ? 0 : p.ntile2 * (p.n2 / p.cg) * 128;
    p.stage_bytes = ((g1 > g2 ? g1 : g2) + 1023) / 1024 * 1024;
  }
  const int bias_bytes = (p.J * 128 + C) * 4;
  // projection mode: a second x buffer when it still leaves a ring of three chunks (narrow models only)
  if (proj && (dp.max_smem - 1024 - 2 * p.KB1 * kFfXBlock - bias_bytes) / p.stage_bytes >= 3) p.nxb = 2;
  p.off_ring = p.nxb * p.KB1 * kFfXBlock;
  int stages = (dp.max_smem - 1024 - p.off_ring - bias_bytes) / p.stage_bytes;
  if (stages > kFfMaxStages) stages = kFfMaxStages;
  CTRLV_CHECK_ARG(stages >= 1, "feedforward: not enough shared memory");
  p.stages = stages;
  p.off_b1 = p.off_ring + stages * p.stage_bytes;
  p.off_b2 = p.off_b1 + p.J * 128 * 4;
  p.b1 = b1;
  p.ep = *ep;
  p.ln = ln; p.ln_eps = ln_eps;
  p.ln_rb = ln_rb; p.ln_rb_ld = ln_rb_ld; p.ln_rb_div = ln_rb_div; p.ln_rb_mod = ln_rb_mod;
  int rc;
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {64, 128};
    rc = encode_tmap_bf16(&p.tmX, x, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)C, (uint64_t)(proj ? proj_N : 2 * p.H)};
    uint64_t strides[1] = {(uint64_t)C * 2};
    uint32_t box[2] = {64, (uint32_t)(128 / p.cg)};
    rc = encode_tmap_bf16(&p.tmW1, W1, 2, dims, strides, box, true);
    if (rc) return rc;
  }
  if (!proj) {
    uint64_t dims[2] = {(uint64_t)p.H, (uint64_t)C};
    uint64_t strides[1] = {(uint64_t)p.H * 2};
    uint32_t box[2] = {64, (uint32_t)(p.n2 / p.cg)};
    rc = encode_tmap_bf16(&p.tmW2, W2, 2, dims, strides, box, true);
    if (rc) return rc;
  } else {
    p.tmW2 = p.tmW1;  // (prefetched, never loaded from)
  }
  const size_t smem = (size_t)p.off_b2 + (size_t)C * 4 + 1024;
  if (p.cg == 1) {
    const int grid = p.tiles < dp.sms ? p.tiles : dp.sms;
    CTRLV_CUDA(launch_pdl(ff_kernel<1>, dim3(grid), dim3(kFfThreads), smem, stream, p));
  } else {
    const int units = (p.tiles + 1) / 2;
    const int pairs = units < dp.sms / 2 ? units : dp.sms / 2;
    CTRLV_CUDA(launch_cluster2(ff_kernel<2>, dim3(2 * pairs), dim3(kFfThreads), smem, stream, p));
  }
  return CTRLV_OK;
}

extern "C" int ctrlv_feedforward(const void* x, int64_t ldx, int32_t M, int32_t C, const void* W1, const float* b1,
                                 const void* W2, const ctrlv_epilogue* ep, void* stream) {
  return feedforward_impl(x, ldx, M, C, 0, 0.f, nullptr, 0, 1, 1, W1, b1, W2, ep, stream);
}

extern "C" int ctrlv_feedforward_ln(const void* x, int64_t ldx, int32_t M, int32_t C, float ln_eps, const float* ln_rowbias,
                                    int32_t ld_ln_rowbias, int32_t ln_rb_div, int32_t ln_rb_mod, const void* W1,
                                    const float* b1, const void* W2, const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ln_eps > 0.f, "feedforward_ln: eps must be positive");
  if (ln_rowbias)
    CTRLV_CHECK_ARG(ld_ln_rowbias >= C && ld_ln_rowbias % 4 == 0 && ln_rb_div > 0 && ln_rb_mod > 0 &&
                        (reinterpret_cast<uintptr_t>(ln_rowbias) & 15) == 0,
                    "feedforward_ln: row-bias table needs ld >= C, ld %% 4 == 0, 16-byte alignment, div / mod > 0");
  return feedforward_impl(x, ldx, M, C, 1, ln_eps, ln_rowbias, ld_ln_rowbias, ln_rowbias ? ln_rb_div : 1,
                          ln_rowbias ? ln_rb_mod : 1, W1, b1, W2, ep, stream);
}

extern "C" int ctrlv_linear_ln(const void* x, int64_t ldx, int32_t M, int32_t K, float ln_eps, const float* ln_rowbias,
                               int32_t ld_ln_rowbias, int32_t ln_rb_div, int32_t ln_rb_mod, const void* W, int32_t N,
                               const ctrlv_epilogue* ep, void* stream) {
  CTRLV_CHECK_ARG(ep != nullptr && ln_eps > 0.f && N > 0 && N % 64 == 0, "linear_ln: eps > 0, N a positive multiple of 64");
  CTRLV_CHECK_ARG(ep->rb_mode == 0 && ep->res1 == nullptr && ep->res2 == nullptr && ep->s_acc == 1.0f,
                  "linear_ln: the epilogue takes a bias and a bf16 out only");
  if (ln_rowbias)
    CTRLV_CHECK_ARG(ld_ln_rowbias >= K && ld_ln_rowbias % 4 == 0 && ln_rb_div > 0 && ln_rb_mod > 0 &&
                        (reinterpret_cast<uintptr_t>(ln_rowbias) & 15) == 0,
                    "linear_ln: row-bias table needs ld >= K, ld %% 4 == 0, 16-byte alignment, div / mod > 0");
  return feedforward_impl(x, ldx, M, K, 1, ln_eps, ln_rowbias, ld_ln_rowbias, ln_rowbias ? ln_rb_div : 1,
                          ln_rowbias ? ln_rb_mod : 1, W, ep->bias, nullptr, ep, stream, N);
}

/* Tuning / test hook: force the CTA-group size of ctrlv_feedforward (0 = automatic). */
extern "C" int ctrlv_feedforward_override(int32_t cta_group) {
  CTRLV_CHECK_ARG(cta_group >= 0 && cta_group <= 2, "feedforward_override: cta_group=%d", cta_group);
  g_ff_force_cg = cta_group;
  return CTRLV_OK;
}
