// Exact unsigned 32-bit division by a launch-time constant (Granlund-Montgomery round-up method):
//   q = (t + ((n - t) >> sh1)) >> sh2  with  t = umulhi(n, mul)
// — 4 instructions instead of the ~25 of a software 32-bit division, which the per-tile index
// decomposition of the implicit GEMM paid six times.  Host + device; checked exhaustively on the CPU
// by tests/test_host.py.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CTRLV_HD __host__ __device__ __forceinline__
#else
#define CTRLV_HD inline
#endif

namespace ctrlv {

struct FastDiv {
  uint32_t d, mul, sh1, sh2;
};

inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) { f.mul = 0; f.sh1 = 0; f.sh2 = 0; return f; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
  f.sh1 = 1;
  f.sh2 = l - 1;
  return f;
}

CTRLV_HD uint32_t fd_div(uint32_t n, const FastDiv& f) {
#if defined(__CUDA_ARCH__)
  const uint32_t t = __umulhi(n, f.mul);
#else
  const uint32_t t = (uint32_t)(((uint64_t)n * f.mul) >> 32);
#endif
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

CTRLV_HD void fd_divmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
  q = fd_div(n, f);
  r = n - q * f.d;
}

}  // namespace ctrlv
