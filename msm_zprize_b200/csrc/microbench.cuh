// Integer-pipe micro-benchmarks: the measured denominators of the IMAD roofline
// (BASELINE.json north_star "Evidence": a pure-IMAD peak on the same B200).
//
// Every variant keeps MB_CHAINS independent dependency chains per thread whose multiplicands are
// data dependent, so ptxas can neither hoist the products nor strength-reduce them; the SASS of
// each variant is checked in profiles/ (one IMAD / IMAD.WIDE / IMAD.HI / IADD3 per counted op).
#pragma once
#include "kernels_common.cuh"
#include "kernels_reduce.cuh"

namespace msm {

constexpr int MB_CHAINS = 8;
constexpr int MB_INNER = 32;

// which: 0 mad.lo.u32 (IMAD), 1 mad.wide.u32 (IMAD.WIDE.U32), 2 mad.lo.cc + madc.hi.cc pairs in
// one carry chain (IMAD.WIDE.U32.X), 5 mad.hi.u32 (IMAD.HI.U32), 8 add.u32 (IADD3)
template <int WHICH>
__global__ void k_mb_imad(uint32_t* out, int iters, uint32_t seed) {
  uint32_t x[MB_CHAINS], y[MB_CHAINS];
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) {
    x[j] = seed + threadIdx.x * 2654435761u + j;
    y[j] = seed * 7 + blockIdx.x + j * 40503u;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < MB_INNER; u++) {
      if (WHICH == 2) {
        // one carry chain across all accumulator pairs, as in a row of the Montgomery product
#pragma unroll
        for (int j = 0; j < MB_CHAINS; j += 2) {
          uint32_t a = x[(j + 2) % MB_CHAINS], b = y[(j + 3) % MB_CHAINS];
          if (j == 0)
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
          else
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(x[j]) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(x[j + 1]) : "r"(a), "r"(b));
        }
#pragma unroll
        for (int j = 0; j < MB_CHAINS; j += 2) {
          uint32_t a = x[(j + 4) % MB_CHAINS], b = x[(j + 5) % MB_CHAINS];
          if (j == 0)
            asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
          else
            asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(y[j]) : "r"(a), "r"(b));
          asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(y[j + 1]) : "r"(a), "r"(b));
        }
      } else {
#pragma unroll
        for (int j = 0; j < MB_CHAINS; j++) {
          if (WHICH == 0) {
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(y[j]), "r"(y[(j + 1) % MB_CHAINS]));
          } else if (WHICH == 5) {
            asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(y[j]), "r"(y[(j + 1) % MB_CHAINS]));
          } else if (WHICH == 1) {
            asm volatile(
                "{\n\t.reg .u64 t;\n\tmov.b64 t, {%0, %1};\n\tmad.wide.u32 t, %2, %3, t;\n\tmov.b64 {%0, %1}, t;\n\t}"
                : "+r"(x[j]), "+r"(y[j])
                : "r"(x[(j + 1) % MB_CHAINS]), "r"(y[(j + 3) % MB_CHAINS]));
          } else {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(x[j]) : "r"(y[j]));
          }
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < MB_CHAINS; j++) acc ^= x[j] ^ y[j];
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

// dependent chain of Montgomery squarings per thread (two chains for ILP, like k_mb_modmul)
template <class F>
__global__ void k_mb_modsqr(uint32_t* out, int iters, uint32_t seed) {
  Fe<F> x = fe_one<F>(), y = fe_one<F>();
  x.v[0] ^= seed + threadIdx.x;
  y.v[1] ^= seed * 5 + blockIdx.x + 7u * threadIdx.x;
  fe_reduce_once(x);
  fe_reduce_once(y);
  for (int it = 0; it < iters; it++) {
    x = fe_sqr(x);
    y = fe_sqr(y);
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < F::N; j++) acc ^= x.v[j] ^ y.v[j];
  if (acc == 0x12345678u) out[0] = acc;
}

// latency of the one-warp tail: a chain of point doublings, quad-cooperative (QUAD) or one lane alone
template <class C, bool QUAD>
__global__ void __launch_bounds__(32) k_mb_dbl_chain(uint32_t* out, int iters, uint32_t seed) {
  typename C::Acc acc = C::zero();
  acc.X.v[0] = seed & 0xffffu;
  acc.Y.v[1] = 5u;
  acc.Z.v[0] = 1u;
#pragma unroll 1
  for (int it = 0; it < iters; it++) acc = QUAD ? C::dblq(acc) : C::dbl(acc);
  uint32_t x = 0;
#pragma unroll
  for (int j = 0; j < C::F::N; j++) x ^= acc.X.v[j] ^ acc.Y.v[j] ^ acc.Z.v[j];
  if (x == 0x12345678u) out[0] = x;
}

}  // namespace msm
