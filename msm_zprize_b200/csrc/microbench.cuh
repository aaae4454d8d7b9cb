// Integer-pipe micro-benchmarks: the measured denominators of the IMAD roofline
// (BASELINE.json north_star "Evidence": a pure-IMAD peak on the same B200).
#pragma once
#include "kernels_common.cuh"

namespace msm {

constexpr int MB_CHAINS = 8;
constexpr int MB_INNER = 64;

// which: 0 mad.lo.u32, 1 mad.wide.u32, 2 mad.lo.cc/madc.hi.cc pairs, 5 mad.hi.u32
template <int WHICH>
__global__ void k_mb_imad(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t x[MB_CHAINS];
  unsigned long long w[MB_CHAINS];
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) {
    x[j] = a + j;
    w[j] = a * 7 + j;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < MB_INNER; u++) {
#pragma unroll
      for (int j = 0; j < MB_CHAINS; j++) {
        if (WHICH == 0) {
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
        } else if (WHICH == 5) {
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
        } else if (WHICH == 1) {
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[j]) : "r"(a), "r"(b));
        } else {
          uint32_t lo = (uint32_t)w[j], hi = (uint32_t)(w[j] >> 32);
          asm volatile(
              "mad.lo.cc.u32 %0, %2, %3, %0;\n\t"
              "madc.hi.u32 %1, %2, %3, %1;"
              : "+r"(lo), "+r"(hi)
              : "r"(a), "r"(b));
          w[j] = ((unsigned long long)hi << 32) | lo;
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) acc ^= x[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
  if (acc == 0x12345678u) out[0] = acc;  // keep the chains alive
}

// dependent chain of Montgomery products per thread
template <class F>
__global__ void k_mb_modmul(uint32_t* out, int iters, uint32_t seed) {
  Fe<F> x = fe_one<F>(), y = fe_one<F>();
  x.v[0] ^= seed + threadIdx.x;
  y.v[1] ^= seed * 5 + blockIdx.x;
  fe_reduce_once(x);
  fe_reduce_once(y);
  for (int it = 0; it < iters; it++) {
    x = fe_mul(x, y);
    y = fe_mul(y, x);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < F::N; j++) acc ^= x.v[j] ^ y.v[j];
  if (acc == 0x12345678u) out[0] = acc;
}

}  // namespace msm
