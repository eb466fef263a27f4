/*
 * ctrlv_b200 — C-ABI of the B200-native Box2Video denoise step.
 *
 * The reference (oooolga/Ctrl-V) has no native layer of its own: its hot path
 *   src/ctrlv/pipelines/pipeline_video_control.py:298-343   (loop body)
 *   src/ctrlv/models/controlnet.py:226-351                  (ControlNetModel.forward)
 *   src/ctrlv/models/unet_spatio_temporal_condition.py:31-171 (UNet forward with residuals)
 * dispatches through diffusers==0.27.2 modules to cuDNN / cuBLAS / SDPA.  The entry points
 * below are what a replacement of that dispatch binds (ctypes stub: INTEGRATION.md).  Each
 * one cites the reference op it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller; the library allocates nothing
 *    on the hot path; every call is asynchronous on the `stream` passed (a cudaStream_t);
 *  - activations are channels-last bf16: a tensor the reference holds as [frames, C, h, w]
 *    is a row-major matrix [frames*h*w, C] here ("rows" = spatial sites, "cols" = channels);
 *  - return value: 0 = ok, negative = error (ctrlv_last_error() gives the message); no C++
 *    exception crosses the boundary.
 */
#ifndef CTRLV_B200_H_
#define CTRLV_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTRLV_OK 0
#define CTRLV_ERR_INVALID (-1)
#define CTRLV_ERR_CUDA (-2)
#define CTRLV_ERR_UNSUPPORTED (-3)

/* Last error message of the calling thread ("" if none). */
const char* ctrlv_last_error(void);
/* Library version string and the SM architecture it was compiled for ("sm_100a"). */
const char* ctrlv_version(void);
/* Number of kernels this library has launched (or recorded into a stream capture) in this process. */
int64_t ctrlv_launch_count(void);
/* 0 if the current device can run the kernels (compute capability 10.x), negative otherwise. */
int ctrlv_device_check(void);

/* ------------------------------------------------------------------------------------------
 * Launch plans (what replaces the loop body pipeline_video_control.py:305-324 for a host that
 * is not Python).  A plan records everything this library is asked to do on the calling thread between
 * ctrlv_plan_create and ctrlv_plan_finish — every kernel launch with its configuration and a copy of its
 * arguments (tensor maps are encoded ONCE, at record time), ctrlv_memset_zero calls, and the fork / join points
 * of a second stream — while executing it normally.  ctrlv_plan_run(plan, stream) then re-issues the whole
 * sequence from C: one call per denoise step (`ctrlv_denoise_step` of SURVEY §8(b) is ctrlv_plan_run on a plan
 * recorded from one ControlNet + UNet + CFG/Euler step), no per-launch host work besides cudaLaunchKernelExC.
 * The caller keeps every buffer the recorded calls referenced alive and at the same address (per-step scalars
 * — sigma, timestep — live in device memory for that reason, see ctrlv_prep_input / ctrlv_cfg_euler).
 * Launches recorded on a stream other than `main_stream` replay on a stream owned by the plan, ordered against
 * the main stream by the recorded ctrlv_plan_fork (side waits for main) / ctrlv_plan_join (main waits for side).
 * ---------------------------------------------------------------------------------------- */
typedef struct ctrlv_plan ctrlv_plan;
int ctrlv_plan_create(void* main_stream, ctrlv_plan** out); /* start recording on the calling thread */
int ctrlv_plan_fork(void);                                  /* no-ops when the thread is not recording */
int ctrlv_plan_join(void);
int ctrlv_plan_finish(ctrlv_plan* plan);                    /* stop recording */
int64_t ctrlv_plan_size(const ctrlv_plan* plan);            /* recorded kernel launches */
int ctrlv_plan_run(ctrlv_plan* plan, void* stream);
int ctrlv_plan_destroy(ctrlv_plan* plan);
/* cudaMemsetAsync(ptr, 0, bytes) that a plan can record (zeroing of GroupNorm statistics tables). */
int ctrlv_memset_zero(void* ptr, int64_t bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused epilogue of every tensor-core contraction (acc = fp32 accumulator of element (m, n)):
 *     v   = acc + bias[n] + rowbias[ridx(m)][n]
 *     v   = geglu ? v_value * gelu_erf(v_gate) : v          (columns interleaved value/gate)
 *     out = s_acc * v + s_res1 * res1[m][n] + s_res2 * res2[m][n]
 * This one form covers, in the reference:  Linear/Conv bias; ResnetBlock2D's
 * `+ time_emb_proj(silu(temb))[:, :, None, None]`; GEGLU (diffusers FeedForward);
 * residual adds; AlphaBlender `a*x_spatial + (1-a)*x_temporal`; the degenerate 1-token
 * cross-attention (a per-sample vector); ControlNet `* conditioning_scale`
 * (controlnet.py:343-344).
 *   rb_mode 0: no rowbias        1: ridx = m / rb_div        2: ridx = (m / rb_div) % rb_mod
 *   rb_mode 3: ridx = ((m / rb_div) * rb_mod + m % rb_mod + rb_off) % rb_B   (diffusers-0.27.2
 *              S-major `time_context` broadcast; rb_off = offset of this process's rows in the
 *              whole batch's site sequence when the CFG branches are sharded over two processes)
 * ---------------------------------------------------------------------------------------- */
typedef struct ctrlv_epilogue {
  const float* bias;    /* [N] fp32 or NULL */
  const float* rowbias; /* [R][ld_rowbias] fp32 or NULL */
  int32_t ld_rowbias;
  int32_t rb_mode, rb_div, rb_mod, rb_B;
  int32_t geglu;     /* 1: N/2 outputs per row */
  float s_acc;       /* scale of the accumulator term, taken literally: set it to 1.0f for a plain
                        product (a zero-initialised struct multiplies the product by 0, which is what
                        conditioning_scale = 0 asks for, controlnet.py:343-344) */
  const void* res1;  /* bf16 [M][ld_res1] or NULL */
  int32_t ld_res1;
  float s_res1;
  const void* res2;  /* bf16 [M][ld_res2] or NULL */
  int32_t ld_res2;
  float s_res2;
  void* out;         /* bf16 [M][ld_out] or NULL */
  int32_t ld_out;
  float* out_f32;    /* fp32 [M][ld_out_f32] or NULL */
  int32_t ld_out_f32;
  int32_t n_store;   /* number of valid output columns (0 = all) */
  /* Optional: statistics for the nn.GroupNorm(32) that consumes `out` (diffusers ResnetBlock2D /
   * TemporalResnetBlock norm1/norm2, TransformerSpatioTemporalModel.norm, conv_norm_out), accumulated
   * by this launch's epilogue so that the norm needs no statistics pass of its own:
   *   gn_sums[r][m / gn_rows_per_unit][(gn_c_off + n) / gn_cg] += (sum, sum of squares) of the stored bf16 value
   * as int64 fixed point (value * 2^16, two's complement) — see ctrlv_groupnorm_apply.  r is one of gn_rep
   * replicas of the table (a power of two; the consumer adds them up): with few units a single table is an L2
   * hot spot for the whole chip.  The caller zeroes gn_sums before the first producer; several producers
   * (skip-concat halves, the four upsample phases) may add into the same table.  Needs gn_cg in {4, 6, >= 7},
   * gn_c_off % 8 == 0, a bf16 `out`, no GEGLU, n_store == 0. */
  void* gn_sums;     /* int64 [gn_rep][gn_units][32][2] or NULL */
  int32_t gn_rows_per_unit, gn_cg, gn_c_off, gn_units, gn_rep;
  int32_t rb_off;    /* rb_mode 3 only (see above); 0 otherwise */
  /* Optional stream-K workspace of the implicit GEMM (NULL: whole tiles only).  Problems whose tiles fill the SMs
   * badly and whose K loop is long (the deepest UNet level: 10 row tiles on 148 SMs, K up to 25600) are cut along
   * K as well: every SM runs the same number of k-blocks; the k-ranges of a shared tile are parked here as fp32
   * partial tiles, and a second, small launch adds them up in contributor order (bit-reproducible, no atomics)
   * and runs this epilogue on the sum.  32-byte aligned, splitk_bytes >= 2 * (number of SMs) * 128 KiB for every
   * problem to qualify (37 MiB on a B200); needs no initialisation.  One workspace per stream: launches that
   * may run concurrently must not share it. */
  void* splitk_ws;
  int64_t splitk_bytes;
} ctrlv_epilogue;

/* One operand source of the implicit GEMM: a channels-last view [Z][Y][X][C] with element
 * strides (sz, sy, sx, 1). */
typedef struct ctrlv_src {
  const void* ptr; /* bf16 */
  int32_t C;       /* channels addressable in this source (multiple of 64) */
  int64_t sx, sy, sz;
} ctrlv_src;

/* One K segment: `nchunk` chunks of 64 channels, starting at channel c0 of source `src`,
 * read at coordinates (x+dx, y+dy, z+dz); out-of-range coordinates read zeros. */
typedef struct ctrlv_seg {
  int32_t src, c0, nchunk, dx, dy, dz;
} ctrlv_seg;

#define CTRLV_MAX_SRC 4
#define CTRLV_MAX_SEG 20

typedef struct ctrlv_igemm_desc {
  int32_t nsrc;
  ctrlv_src src[CTRLV_MAX_SRC];
  int32_t X, Y, Z; /* output rows m = (z*Y + y)*X + x */
  int32_t nseg;
  ctrlv_seg seg[CTRLV_MAX_SEG];
  const void* W;   /* bf16 [N][K], K = 64 * sum(nchunk), K contiguous */
  int32_t N, K;
  int32_t bn;      /* n-tile override (0 = auto) */
  /* Optional output-row remap (all zero = identity): the row of (x, y, z) in out / res1 / res2 /
   * rowbias indexing becomes (z*out_Y + y*out_mul_y + out_off_y)*out_X + x*out_mul_x + out_off_x.
   * Used to scatter the four parity phases of a fused nearest-2x-upsample + 3x3 conv. */
  int32_t out_mul_x, out_mul_y, out_off_x, out_off_y, out_X, out_Y;
  ctrlv_epilogue ep;
} ctrlv_igemm_desc;

/* The generic tcgen05 implicit-GEMM contraction (TMA-fed, TMEM accumulators, persistent).
 * Everything below that is a convolution or a Linear is a thin wrapper over it. */
int ctrlv_igemm(const ctrlv_igemm_desc* desc, void* stream);

/* Introspection (no CUDA calls): the row box (bx, by, bz), n-tile width and cta_group (1 or 2) the
 * launcher picks for `desc` on a device with `num_sms` SMs — lets host tests pin the tiling heuristics. */
int ctrlv_igemm_plan(const ctrlv_igemm_desc* desc, int32_t num_sms, int32_t* box_xyz, int32_t* bn,
                     int32_t* cta_group);

/* Tuning hook for tile sweeps (process-global, not for production use): force the n-tile width, the
 * cta_group and the pipeline depth of every following ctrlv_igemm launch; 0 = the launcher's heuristic. */
int ctrlv_igemm_override(int32_t bn, int32_t cta_group, int32_t stages);
/* Tuning / test hook (process-global): stream-K scheduling of launches that carry a workspace
 * (ctrlv_epilogue.splitk_ws): 0 = the launcher's heuristic, 1 = never, 2 = whenever the problem allows it. */
int ctrlv_igemm_streamk(int32_t mode);

/* nn.Linear (diffusers Attention.to_q/k/v/to_out, FeedForward, proj_in/out, 1x1 convs incl. the
 * ControlNet zero-convs controlnet.py:148-185,331-339):  out[M][N] = A[M][K] * W[N][K]^T. */
int ctrlv_linear(const void* A, int64_t lda, int32_t M, int32_t K, const void* W, int32_t N,
                 const ctrlv_epilogue* ep, void* stream);

/* diffusers FeedForward(dim = C, activation "geglu", mult 4) of BasicTransformerBlock.ff and
 * TemporalBasicTransformerBlock.ff_in / .ff as ONE launch for C <= 320 (level 0 of the SVD UNet):
 *     out = s_acc * (GEGLU(x W1^T + b1) W2^T + bias + rowbias[ridx(m)]) + s_res1 * res1 + s_res2 * res2
 * x [M][ldx] bf16 (C columns used); W1 [8C][C] with the GEGLU (value, gate) rows interleaved, b1 [8C] fp32
 * likewise; W2 [C][4C]; `ep` is the output epilogue (bias = b2 [C], rowbias, s_acc, res1, res2, bf16 out with
 * 32-byte aligned rows).  The [M][4C] intermediate stays in tensor memory (see csrc/ff.cu).  Same result as
 * ctrlv_linear(geglu) followed by ctrlv_linear up to the bf16 rounding of the intermediate (identical here:
 * the hidden activations are rounded to bf16 before the second contraction in both forms). */
int ctrlv_feedforward(const void* x, int64_t ldx, int32_t M, int32_t C, const void* W1, const float* b1,
                      const void* W2, const ctrlv_epilogue* ep, void* stream);
/* The same FeedForward with the nn.LayerNorm(C, ln_eps) in front of it (BasicTransformerBlock.norm3,
 * TemporalBasicTransformerBlock.norm_in / .norm3) in the same launch: x holds the rows BEFORE the norm, every
 * 128-row tile is normalised in shared memory (fp32 two-pass statistics, one rounding to bf16 — the arithmetic of
 * ctrlv_layernorm with gamma == beta == NULL) before the first contraction reads it; the norm's affine part is
 * folded into W1 / b1 by the caller.  ln_rowbias (or NULL): fp32 [rows][ld_ln_rowbias] added to x before the
 * statistics, row m takes ln_rowbias[(m / ln_rb_div) % ln_rb_mod] (the frame-position embedding in front of norm_in). */
int ctrlv_feedforward_ln(const void* x, int64_t ldx, int32_t M, int32_t C, float ln_eps, const float* ln_rowbias,
                         int32_t ld_ln_rowbias, int32_t ln_rb_div, int32_t ln_rb_mod, const void* W1, const float* b1,
                         const void* W2, const ctrlv_epilogue* ep, void* stream);
/* nn.LayerNorm(K, ln_eps) + nn.Linear(K, N) in one launch for K <= 320 (BasicTransformerBlock.norm1 -> attn1.to_q /
 * to_k / to_v fused, TemporalBasicTransformerBlock.norm1 -> attn1): the machinery of ctrlv_feedforward_ln without the
 * GEGLU and the second contraction — the x tile is normalised in shared memory, out[m][n] = LN(x)[m] . W[n] + bias[n]
 * leaves the launch chunk by chunk (128 columns at a time) as bf16.  W [N][K] bf16 (the norm's affine part folded in
 * by the caller), N % 64 == 0; `ep` carries the bias (or NULL) and the bf16 out (32-byte aligned rows), nothing else. */
int ctrlv_linear_ln(const void* x, int64_t ldx, int32_t M, int32_t K, float ln_eps, const float* ln_rowbias,
                    int32_t ld_ln_rowbias, int32_t ln_rb_div, int32_t ln_rb_mod, const void* W, int32_t N,
                    const ctrlv_epilogue* ep, void* stream);
/* Tuning / test hook (process-global): force single CTAs (1) or CTA pairs (2) in ctrlv_feedforward; 0 = automatic. */
int ctrlv_feedforward_override(int32_t cta_group);

/* nn.Conv2d 3x3, padding 1, stride 1 or 2 (ResnetBlock2D.conv1/conv2, Downsample2D,
 * Upsample2D.conv, conv_in, conv_out) on channels-last frames [frames][H][W][C].
 *  - up to two input sources (src1 may be NULL): channel concat of the UNet skip connection
 *    (torch.cat in the up blocks) without materialising it;
 *  - optional fused 1x1 `conv_shortcut` over up to two raw sources appended to the K loop;
 *    weights W[N][9*(C0+C1) + (SC0+SC1)], tap-major then channel. */
int ctrlv_conv3x3(const void* src0, int32_t C0, const void* src1, int32_t C1, int32_t frames,
                  int32_t H, int32_t Wd, int32_t stride, const void* sc0, int32_t SC0,
                  const void* sc1, int32_t SC1, const void* W, int32_t N,
                  const ctrlv_epilogue* ep, void* stream);

/* nn.Conv3d kernel (3,1,1), padding (1,0,0) of TemporalResnetBlock on [B][T][HW][C]
 * (zero halo per clip); weights W[N][3*C], tap-major. */
int ctrlv_conv_t3(const void* src, int32_t C, int32_t B, int32_t T, int32_t HW, const void* W,
                  int32_t N, const ctrlv_epilogue* ep, void* stream);

/* nn.GroupNorm(32, C0+C1, eps) [+ SiLU] over `n_units` statistics units of `rows_per_unit`
 * consecutive rows each (spatial norm: unit = one frame; TemporalResnetBlock norm: unit = one
 * clip, i.e. statistics across frames).  Reads one or two sources (channel concat), writes the
 * normalised (and activated) concat as bf16 [rows][C0+C1].
 * `workspace` : fp32, at least ctrlv_groupnorm_workspace(n_units) bytes. */
int64_t ctrlv_groupnorm_workspace(int32_t n_units);
int ctrlv_groupnorm(const void* src0, int32_t C0, const void* src1, int32_t C1, int32_t n_units,
                    int32_t rows_per_unit, const float* gamma, const float* beta, float eps,
                    int32_t silu, void* out, void* workspace, void* stream);

/* The normalise(+SiLU) half of ctrlv_groupnorm for an input whose statistics were accumulated by its
 * producers (ctrlv_epilogue.gn_sums / ctrlv_axpby_gn): sums = int64 [n_rep][n_units][32][2] fixed point
 * (sum, sum of squares) * 2^16, the n_rep replicas being added up.  Same arithmetic as ctrlv_groupnorm from the statistics on. */
int ctrlv_groupnorm_apply(const void* src0, int32_t C0, const void* src1, int32_t C1, int32_t n_units,
                          int32_t rows_per_unit, const float* gamma, const float* beta, float eps,
                          int32_t silu, void* out, const void* sums, int32_t n_rep, void* stream);

/* out = a*x + b*y on bf16 rows [rows][C] (the ControlNet residual add into a skip connection,
 * unet_spatio_temporal_condition.py:119-127,136-137) that also accumulates the GroupNorm statistics of
 * `out` for its consumer, exactly like ctrlv_epilogue.gn_sums: group (c_off + c) / cg of unit row /
 * rows_per_unit.  y == out == NULL: only the statistics of x (a skip connection used as is: the
 * same reduction, so a zero residual gives bit-identical results to no residual). */
int ctrlv_axpby_gn(const void* x, const void* y, float a, float b, int64_t rows, int32_t C, void* out,
                   void* gn_sums, int32_t rows_per_unit, int32_t cg, int32_t c_off, int32_t n_rep, void* stream);

/* nn.LayerNorm(C, eps) over rows, with an optional fp32 row-bias added first
 * (x + rowbias[(m / rb_div) % rb_mod]): the frame-position embedding of
 * TransformerSpatioTemporalModel.  gamma == beta == NULL: plain normalisation (the affine part
 * folded into the weights of the Linear that consumes the result). */
int ctrlv_layernorm(const void* x, int64_t ldx, int32_t M, int32_t C, const float* gamma,
                    const float* beta, float eps, const float* rowbias, int32_t ld_rowbias,
                    int32_t rb_div, int32_t rb_mod, void* out, void* stream);

/* Self-attention core (F.scaled_dot_product_attention in AttnProcessor2_0), head_dim 64, no
 * mask, on a fused projection buffer qkv[rows][3*C] (q | k | v, heads of 64 channels).
 *  spatial : rows = frames*S, sequences are the S sites of one frame;
 *  temporal: rows = B*T*S, sequences are the T frames of one site (row stride S between
 *            sequence elements) — no permute copy.
 * out: bf16 [rows][C], 32-byte aligned (rows are written with 256-bit stores); qkv 16-byte aligned. */
int ctrlv_attn_spatial(const void* qkv, int32_t frames, int32_t S, int32_t heads, float scale,
                       void* out, void* stream);
int ctrlv_attn_temporal(const void* qkv, int32_t B, int32_t T, int32_t S, int32_t heads,
                        float scale, void* out, void* stream);

/* Cross-attention core for a context of L > 1 tokens (diffusers Attention with encoder_hidden_states
 * [batch, L, D]; controlnet.py:230,244-245): q [M][ldq] bf16 (heads of 64 channels, already projected by to_q),
 * kv [n_ctx*L][ldkv] bf16 = to_k(ctx) | to_v(ctx) side by side (K in columns [0, C), V in [C, 2C)), L <= 256.
 * Row m attends to context ridx(m), ridx as in ctrlv_epilogue.rb_mode (1: m / div, 2: (m / div) % mod,
 * 3: the diffusers-0.27.2 S-major time_context order).  out [M][C] bf16.  L = 1 needs no kernel: the
 * softmax over one key is 1 (the pipelines' case, folded into a per-sample vector by the host). */
int ctrlv_cross_attn(const void* q, int64_t ldq, const void* kv, int64_t ldkv, int32_t M, int32_t heads, int32_t L,
                     int32_t n_ctx, float scale, int32_t ctx_mode, int32_t ctx_div, int32_t ctx_mod, int32_t ctx_B,
                     void* out, void* stream);

/* Small-M dense layers on CUDA cores (embedding MLPs, time_emb_proj, 1-token cross-attention
 * value path): y[M][N] = act_in(x)[M][K] * W[N][K]^T + b, M <= 32; x, y fp32; W bf16.
 * act_in: 0 none, 1 SiLU.  act_out: 0 none, 1 SiLU.  accumulate: 1 adds the result to y
 * (emb = time_embedding(...) + add_embedding(...), controlnet.py:277-283). */
int ctrlv_small_linear(const float* x, int32_t M, int32_t K, const void* W, const float* bias,
                       int32_t N, int32_t act_in, int32_t act_out, int32_t accumulate, float* y,
                       void* stream);

/* diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[n][dim] =
 * [cos(t*f_k) | sin(t*f_k)], f_k = exp(-ln(10000) k / (dim/2)); t fp32 [n]. `round_bf16`
 * rounds the result through bf16 (the reference casts to the model dtype). */
int ctrlv_sinusoid(const float* t, int32_t n, int32_t dim, int32_t round_bf16, float* out,
                   void* stream);

/* Loop-body glue of pipeline_video_control.py:300-304: builds the channels-last, 64-channel
 * padded model input  [2B or B][T][h][w][64] = [latents/sqrt(sigma^2+1) | image_latents |
 * control_cond | 0...]  from NCHW fp32 latents [B][T][4][h][w] and NCHW fp32 conditioning.  sigma is read from device
 * memory so that one captured CUDA graph serves every step of the schedule. */
int ctrlv_prep_input(const float* latents, const float* image_latents, const float* control_cond,
                     int32_t B, int32_t cfg, int32_t T, int32_t h, int32_t w,
                     const float* sigma_dev /* device: [sigma] */, void* out, void* stream);

/* CFG combine + EulerDiscreteScheduler.step (v-prediction), pipeline_video_control.py:327-332:
 * latents (fp32 NCHW [B][T][4][h][w]) updated in place from the model output
 * noise [2B or B][T*h*w][ld_noise] fp32 channels-last (first 4 columns). guidance: [T] fp32. */
int ctrlv_cfg_euler(float* latents, const float* noise, int32_t ld_noise, int32_t B, int32_t cfg,
                    int32_t T, int32_t h, int32_t w, const float* guidance,
                    const float* sigma_dev /* device: [sigma, sigma_next] */, int32_t round_bf16,
                    void* stream);

/* Elementwise / layout helpers. */
int ctrlv_upsample2x(const void* src, int32_t frames, int32_t H, int32_t Wd, int32_t C, void* out,
                     void* stream); /* F.interpolate(scale_factor=2, mode="nearest"), NHWC */
int ctrlv_axpby(const void* x, const void* y, float a, float b, int64_t n, void* out,
                void* stream); /* out = a*x + b*y, bf16 */
/* [frames][C][HW] (fp32 or bf16) -> columns c_off..c_off+C of out[frames*HW][Cpad] (bf16); other
 * columns are left untouched (zero-initialise the buffer once). */
int ctrlv_nchw_to_nhwc(const void* src, int32_t src_is_f32, int32_t frames, int32_t C, int32_t HW,
                       int32_t Cpad, int32_t c_off, void* out, void* stream);
int ctrlv_nhwc_to_nchw(const void* src, int32_t src_is_f32, int64_t ld, int32_t frames, int32_t C,
                       int32_t HW, int32_t out_is_f32, void* out, void* stream);

/* diffusers Downsample2D(padding=0) of the VAE encoder: F.pad(x, (0,1,0,1)) then Conv2d 3x3 stride 2
 * without padding, on channels-last frames; W [N][9*C] tap-major. */
int ctrlv_conv3x3_s2_pad01(const void* src, int32_t C, int32_t frames, int32_t H, int32_t Wd, const void* W,
                           int32_t N, const ctrlv_epilogue* ep, void* stream);

/* diffusers Upsample2D: F.interpolate(scale_factor=2, mode="nearest") followed by Conv2d 3x3 pad 1, fused as
 * four 2x2 phase convolutions of the LOW-resolution frame (2.25x fewer MACs, no upsampled intermediate).
 * src [frames][H][Wd][C]; out rows [frames][2H][2Wd][N]; Wp [4][N][4*C]: phase p = py*2+px holds, for the
 * 2x2 patch offsets (dy, dx) in increasing order, the sums of the 3x3 taps that fall on that source pixel
 * (py = 0: rows {ky=0 | ky=1,2} at dy = {-1, 0};  py = 1: {ky=0,1 | ky=2} at dy = {0, +1};  same in x).
 * Bias-only epilogue. */
int ctrlv_upsample2x_conv3x3(const void* src, int32_t C, int32_t frames, int32_t H, int32_t Wd, const void* Wp,
                             int32_t N, const ctrlv_epilogue* ep, void* stream);

/* ---- temporal VAE glue (diffusers AutoencoderKLTemporalDecoder, reached from
 * pipeline_video_control.py:84,235,346-347) -------------------------------------------------- */

/* probs[m][n] = softmax_n(scale * scores[m][n]) (fp32 in, bf16 out): the softmax between the two
 * GEMMs of the VAE mid-block attention (one head of width 512, F.scaled_dot_product_attention). */
int ctrlv_softmax_rows(const float* scores, int64_t ld_scores, int32_t M, int32_t N, float scale,
                       void* probs, int64_t ld_probs, void* stream);

/* TemporalDecoder.time_conv_out = Conv3d(C, C, (3,1,1), padding (1,0,0)), C <= 4, fused with the
 * channels-last -> NCHW conversion: x [B][T][HW][ld] fp32 (first C columns), w [C][C][3] fp32,
 * out [B*T][C][HW] fp32. */
int ctrlv_time_conv_out(const float* x, int32_t ld, int32_t B, int32_t T, int32_t HW, int32_t C,
                        const float* w, const float* bias, float* out, void* stream);

/* ---- image-conditioning prologue (`_encode_image` of the diffusers SVD pipeline
 * reached from pipeline_video_control.py:220; `_resize_with_antialiasing` as restated in
 * src/ctrlv/bbox_generator_baseline/utils/image_encoder.py:184-290, used by
 * src/ctrlv/utils/util.py:97-125) ------------------------------------------------------------- */

/* One pass of the separable Gaussian blur with reflect padding ((ks-1)/2 in front) over
 * [planes][H][W] fp32; axis 0 = along x, 1 = along y; taps[ks] already normalised.  src != dst. */
int ctrlv_blur1d_reflect(const float* src, int32_t planes, int32_t H, int32_t W, int32_t axis,
                         const float* taps, int32_t ks, float* dst, void* stream);

/* F.interpolate(mode="bicubic", align_corners=True) of [planes][H][W] -> [planes][Ho][Wo], fp32. */
int ctrlv_resize_bicubic_ac(const float* src, int32_t planes, int32_t H, int32_t W, int32_t Ho,
                            int32_t Wo, float* dst, void* stream);

/* CLIP input normalisation + patch gather: img [B][C][H][W] fp32 -> rows [B*(H/P)*(W/P)][Kpad] bf16,
 * column c*P*P + iy*P + ix = (u - mean[c]) / std[c], u = a*img + s (clamped to [0,1] if clamp01);
 * columns >= C*P*P are zero.  The patch-embedding Conv2d(kernel = stride = P) is then ctrlv_linear. */
int ctrlv_clip_patchify(const float* img, int32_t B, int32_t C, int32_t H, int32_t W, int32_t P, float a,
                        float s, int32_t clamp01, const float* mean, const float* stdv, int32_t Kpad,
                        void* rows, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTRLV_B200_H_ */
