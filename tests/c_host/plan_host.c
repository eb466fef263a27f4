/* A host that is not Python: drives libctrlv_b200.so through include/ctrlv_b200.h alone (plain C, cudart for
 * memory).  Records a small launch plan — a GroupNorm-statistics producer Linear, the GroupNorm that consumes
 * its statistics, a fused FeedForward — replays it twice with ctrlv_plan_run and checks that the replays
 * reproduce the directly issued result bit for bit.  Built and run by tests/test_gpu_c_host.py. */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ctrlv_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, ctrlv_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint16_t bf16(float f) { uint32_t u; memcpy(&u, &f, 4); u += 0x7fffu + ((u >> 16) & 1u); return (uint16_t)(u >> 16); }
static float frand(void) { return (float)rand() / (float)RAND_MAX - 0.5f; }

int main(void) {
  const int M = 2048, C = 128, units = 8, rows = M / units;
  if (ctrlv_device_check() != 0) { fprintf(stderr, "no sm_100 device: %s\n", ctrlv_last_error()); return 2; }
  srand(1);
  uint16_t* h = (uint16_t*)malloc((size_t)M * C * 2);
  void *x, *w, *y, *yn, *w1, *w2, *out, *out_ref, *sums;
  float *bias, *gamma, *beta, *b1, *b2;
  CU(cudaMalloc(&x, (size_t)M * C * 2)); CU(cudaMalloc(&w, (size_t)C * C * 2)); CU(cudaMalloc(&y, (size_t)M * C * 2));
  CU(cudaMalloc(&yn, (size_t)M * C * 2)); CU(cudaMalloc(&w1, (size_t)8 * C * C * 2)); CU(cudaMalloc(&w2, (size_t)4 * C * C * 2));
  CU(cudaMalloc(&out, (size_t)M * C * 2)); CU(cudaMalloc(&out_ref, (size_t)M * C * 2));
  const int rep = 16; /* replicas of the statistics table: power of two, >= 128 / units rows in total */
  CU(cudaMalloc(&sums, (size_t)rep * units * 64 * 8));
  CU(cudaMalloc((void**)&bias, C * 4)); CU(cudaMalloc((void**)&gamma, C * 4)); CU(cudaMalloc((void**)&beta, C * 4));
  CU(cudaMalloc((void**)&b1, 8 * C * 4)); CU(cudaMalloc((void**)&b2, C * 4));
  for (int i = 0; i < M * C; ++i) h[i] = bf16(frand() * 2.f);
  CU(cudaMemcpy(x, h, (size_t)M * C * 2, cudaMemcpyHostToDevice));
  uint16_t* hw = (uint16_t*)malloc((size_t)8 * C * C * 2);
  for (int i = 0; i < C * C; ++i) hw[i] = bf16(frand() * 0.2f);
  CU(cudaMemcpy(w, hw, (size_t)C * C * 2, cudaMemcpyHostToDevice));
  for (int i = 0; i < 8 * C * C; ++i) hw[i] = bf16(frand() * 0.2f);
  CU(cudaMemcpy(w1, hw, (size_t)8 * C * C * 2, cudaMemcpyHostToDevice));
  for (int i = 0; i < 4 * C * C; ++i) hw[i] = bf16(frand() * 0.1f);
  CU(cudaMemcpy(w2, hw, (size_t)4 * C * C * 2, cudaMemcpyHostToDevice));
  float hf[8 * 128];
  for (int i = 0; i < 8 * C; ++i) hf[i] = frand();
  CU(cudaMemcpy(bias, hf, C * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(beta, hf + C, C * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(b2, hf + 2 * C, C * 4, cudaMemcpyHostToDevice)); CU(cudaMemcpy(b1, hf, 8 * C * 4, cudaMemcpyHostToDevice));
  for (int i = 0; i < C; ++i) hf[i] = 1.0f + 0.1f * frand();
  CU(cudaMemcpy(gamma, hf, C * 4, cudaMemcpyHostToDevice));

  cudaStream_t s;
  CU(cudaStreamCreate(&s));
  ctrlv_epilogue e1, e2;
  memset(&e1, 0, sizeof(e1));
  e1.bias = bias; e1.s_acc = 1.0f; e1.out = y; e1.ld_out = C;
  e1.gn_sums = sums; e1.gn_rows_per_unit = rows; e1.gn_cg = C / 32; e1.gn_c_off = 0; e1.gn_units = units; e1.gn_rep = rep;
  memset(&e2, 0, sizeof(e2));
  e2.bias = b2; e2.s_acc = 1.0f; e2.res1 = x; e2.ld_res1 = C; e2.s_res1 = 1.0f; e2.out = out; e2.ld_out = C;

  /* the step, issued directly and recorded at the same time */
  ctrlv_plan* plan = NULL;
  CK(ctrlv_plan_create(s, &plan));
  CK(ctrlv_memset_zero(sums, (int64_t)rep * units * 64 * 8, s));
  CK(ctrlv_linear(x, C, M, C, w, C, &e1, s));                                   /* y = x W^T + b, + GroupNorm statistics */
  CK(ctrlv_groupnorm_apply(y, C, NULL, 0, units, rows, gamma, beta, 1e-5f, 1, yn, sums, rep, s)); /* SiLU(GN(y)) */
  CK(ctrlv_feedforward(yn, C, M, C, w1, b1, w2, &e2, s));                       /* out = FF(yn) + x */
  CK(ctrlv_plan_finish(plan));
  CU(cudaStreamSynchronize(s));
  CU(cudaMemcpyAsync(out_ref, out, (size_t)M * C * 2, cudaMemcpyDeviceToDevice, s));
  printf("recorded %lld launches\n", (long long)ctrlv_plan_size(plan));
  if (ctrlv_plan_size(plan) != 3) { fprintf(stderr, "expected 3 recorded launches\n"); return 1; }

  uint16_t* r0 = (uint16_t*)malloc((size_t)M * C * 2);
  uint16_t* r1 = (uint16_t*)malloc((size_t)M * C * 2);
  CU(cudaMemcpyAsync(r0, out_ref, (size_t)M * C * 2, cudaMemcpyDeviceToHost, s));
  for (int it = 0; it < 2; ++it) {
    CU(cudaMemsetAsync(out, 0xff, (size_t)M * C * 2, s));   /* make sure the replay really rewrites everything */
    CU(cudaMemsetAsync(y, 0xff, (size_t)M * C * 2, s));
    CK(ctrlv_plan_run(plan, s));
    CU(cudaMemcpyAsync(r1, out, (size_t)M * C * 2, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (memcmp(r0, r1, (size_t)M * C * 2) != 0) { fprintf(stderr, "replay %d differs from the direct result\n", it); return 1; }
  }
  int finite = 1;
  for (int i = 0; i < M * C; ++i) if ((r0[i] & 0x7f80u) == 0x7f80u) finite = 0;
  if (!finite) { fprintf(stderr, "non-finite output\n"); return 1; }
  CK(ctrlv_plan_destroy(plan));
  printf("plan replay == direct result (%d x %d bf16), OK\n", M, C);
  return 0;
}
