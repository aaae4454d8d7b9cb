// Prime-field arithmetic on 32-bit limbs held in registers (Montgomery form, R = 2^(32 N)).
//
// Replaces the reference's runtime-generated wasm field module:
//   multiply / square            src/wasm/multiply-montgomery.ts:58-215  (29-bit limbs, i64 locals)
//   add / subtract / reduce      src/wasm/field-arithmetic.ts:32-166
//   inverse                      src/wasm/inverse.ts:42-218
// The representation differs on purpose (only the normalised affine output is
// compared): values are kept CANONICAL in [0, p) so that equality tests are limb compares.
//
// Multiplication is operand scanning with the products of even and odd limbs of `a` kept in two
// separate accumulators (one word apart), so that every 32x32->64 product lands on an aligned
// register pair and each row is two independent `mad.lo.cc / madc.hi.cc` carry chains -- ptxas
// fuses each lo/hi pair into one IMAD.WIDE.U32(.X) with the carry in a predicate register.
//
// The same source compiles for the host (carry flag emulated) so that the math core can be unit
// tested without a GPU (tests/test_hostmath.py); kernels never run on the host.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MSM_HD __host__ __device__ __forceinline__
#else
#define MSM_HD inline
#endif

namespace msm {

// ------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
__device__ __forceinline__ uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t mul_lo(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// full 32x32 -> 64 product as ONE wide multiply (separate mul.lo / mul.hi stay two instructions)
__device__ __forceinline__ void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ uint32_t mul_hi(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
#else
// Host emulation of the PTX condition-code carry flag (unit tests only).
inline uint32_t& msm_cf() {
  static thread_local uint32_t cf = 0;
  return cf;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b;
  msm_cf() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + msm_cf();
  msm_cf() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + msm_cf()); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b;
  msm_cf() = (uint32_t)((t >> 32) & 1);  // borrow
  return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - msm_cf();
  msm_cf() = (uint32_t)((t >> 32) & 1);
  return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - msm_cf()); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mul_wide(uint32_t a, uint32_t b, uint32_t& lo, uint32_t& hi) {
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
}
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
#endif

// ------------------------------------------------------------------------------------------
// field element
// ------------------------------------------------------------------------------------------
template <class F>
struct Fe {
  uint32_t v[F::N];
};

template <class F>
MSM_HD bool fe_is_zero(const Fe<F>& a) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < F::N; i++) acc |= a.v[i];
  return acc == 0;
}

template <class F>
MSM_HD bool fe_eq(const Fe<F>& a, const Fe<F>& b) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < F::N; i++) acc |= a.v[i] ^ b.v[i];
  return acc == 0;
}

template <class F>
MSM_HD Fe<F> fe_zero() {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = 0;
  return r;
}

template <class F>
MSM_HD Fe<F> fe_one() {  // Montgomery form of 1
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = F::ONE(i);
  return r;
}

// r = a - p if a >= p else a      (a < 2p)
template <class F>
MSM_HD void fe_reduce_once(Fe<F>& a) {
  uint32_t t[F::N];
  t[0] = sub_cc(a.v[0], F::P(0));
#pragma unroll
  for (int i = 1; i < F::N; i++) t[i] = subc_cc(a.v[i], F::P(i));
  uint32_t borrow = subc(0u, 0u);  // 0xffffffff if a < p
#pragma unroll
  for (int i = 0; i < F::N; i++) a.v[i] = borrow ? a.v[i] : t[i];
}

// (a + b) mod p, canonical inputs    (src/wasm/field-arithmetic.ts:32-63 `add`)
template <class F>
MSM_HD Fe<F> fe_add(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < F::N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[F::N - 1] = addc(a.v[F::N - 1], b.v[F::N - 1]);  // 2p < 2^(32N): no carry out
  fe_reduce_once(r);
  return r;
}

// (a - b) mod p, canonical inputs    (src/wasm/field-arithmetic.ts:65-100 `subtract`)
template <class F>
MSM_HD Fe<F> fe_sub(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> r;
  r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < F::N; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0u, 0u);  // all ones if a < b
  r.v[0] = add_cc(r.v[0], F::P(0) & borrow);
#pragma unroll
  for (int i = 1; i < F::N - 1; i++) r.v[i] = addc_cc(r.v[i], F::P(i) & borrow);
  r.v[F::N - 1] = addc(r.v[F::N - 1], F::P(F::N - 1) & borrow);
  return r;
}

template <class F>
MSM_HD Fe<F> fe_neg(const Fe<F>& a) {
  return fe_sub(fe_zero<F>(), a);
}

template <class F>
MSM_HD Fe<F> fe_dbl(const Fe<F>& a) {
  return fe_add(a, a);
}

template <class F>
MSM_HD Fe<F> fe_select(bool c, const Fe<F>& a, const Fe<F>& b) {  // c ? a : b
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}

// ------------------------------------------------------------------------------------------
// Montgomery product  a*b*2^(-32N) mod p, canonical in/out.          (THE hot function)
// (replaces src/wasm/multiply-montgomery.ts:58-136)
//
// Measured on B200 (tools/microbench.py, profiles/): IMAD.WIDE.U32 issues at 9.1e12/s chip-wide
// (one warp instruction per 4 cycles per SM sub-partition), with or without carry in/out, and
// IMAD.HI at the same rate -- so one IMAD.WIDE per 32x32->64 limb product is the cheapest form
// and 9.1e12 limb products/s is the roofline.  This product issues 2N^2 + N IMAD-pipe
// instructions (N = 12: 279 IMAD.WIDE + 14 IMAD) and runs at 30.0e9/s = 98% of that bound.
// A 29-bit-limb variant with carry-free 64-bit accumulation (the reference's own scheme) was
// measured slower (23.2e9/s): it needs 13 limbs (328 IMAD.WIDE) plus ~390 ALU instructions.
//
// State between rows: two accumulators.  X is word aligned (word w has weight 2^(32w), N+1
// words), Y sits one word higher (word w has weight 2^(32(w+1)), N words); T = X + Y*2^32.
// Row i adds a*b_i (even limbs of a into X, odd limbs into Y), then m*p with m = -T_0/p mod 2^32
// the same way, which clears X_0.  Dividing by 2^32 swaps the roles: new X = Y, new Y_w = X_(w+2)
// and the stray word X_1 is added into new X_0 with its carry entering the new Y chain.
// Bounds (a, b < p): T < 2p < 2^(32N) at every row end, so Y never carries out of word N-1.
// ------------------------------------------------------------------------------------------
template <class F, bool FIRST>
MSM_HD void mont_row(uint32_t* U, uint32_t* V, const uint32_t* a, uint32_t bi) {
  constexpr int N = F::N;
  // On entry (not FIRST): U = previous X (N+1 words), V = previous Y (N words, V[N] unused).
  // On exit: V = X (N+1 words), U = Y (N words)  -- before the division by 2^32.
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      mul_wide(a[j], bi, V[j], V[j + 1]);
      mul_wide(a[j + 1], bi, U[j], U[j + 1]);
    }
    V[N] = 0;
  } else {
    // stray word, then the odd-limb chain reading U two words up
    V[0] = add_cc(V[0], U[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
      U[j] = madc_lo_cc(a[j + 1], bi, U[j + 2]);
      U[j + 1] = madc_hi_cc(a[j + 1], bi, (j + 3 <= N) ? U[j + 3] : 0u);
    }
    U[N - 2] = madc_lo_cc(a[N - 1], bi, U[N]);
    U[N - 1] = madc_hi(a[N - 1], bi, 0u);
    // even-limb chain
    V[0] = mad_lo_cc(a[0], bi, V[0]);
    V[1] = madc_hi_cc(a[0], bi, V[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
      V[j] = madc_lo_cc(a[j], bi, V[j]);
      V[j + 1] = madc_hi_cc(a[j], bi, V[j + 1]);
    }
    V[N] = addc(0u, 0u);
  }
  uint32_t m = mul_lo(V[0], F::M0v());
  // m*p, odd limbs of p into Y
  U[0] = mad_lo_cc(F::P(1), m, U[0]);
  U[1] = madc_hi_cc(F::P(1), m, U[1]);
#pragma unroll
  for (int j = 2; j < N - 2; j += 2) {
    U[j] = madc_lo_cc(F::P(j + 1), m, U[j]);
    U[j + 1] = madc_hi_cc(F::P(j + 1), m, U[j + 1]);
  }
  if (N > 2) {
    U[N - 2] = madc_lo_cc(F::P(N - 1), m, U[N - 2]);
    U[N - 1] = madc_hi(F::P(N - 1), m, U[N - 1]);
  }
  // even limbs of p into X
  V[0] = mad_lo_cc(F::P(0), m, V[0]);
  V[1] = madc_hi_cc(F::P(0), m, V[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    V[j] = madc_lo_cc(F::P(j), m, V[j]);
    V[j + 1] = madc_hi_cc(F::P(j), m, V[j + 1]);
  }
  V[N] = addc(V[N], 0u);
}

template <class F>
MSM_HD Fe<F> fe_mul(const Fe<F>& a, const Fe<F>& b) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t A0[N + 1], A1[N + 1];
  A0[N] = 0;
  A1[N] = 0;
  mont_row<F, true>(A0, A1, a.v, b.v[0]);  // X = A1, Y = A0
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    mont_row<F, false>(A1, A0, a.v, b.v[i]);  // prev X = A1, prev Y = A0 -> X = A0, Y = A1
    if (i + 1 < N) mont_row<F, false>(A0, A1, a.v, b.v[i + 1]);
  }
  // N even: the last row left X = A0 (N+1 words), Y = A1.   T = Y + (X >> 32)
  Fe<F> r;
  r.v[0] = add_cc(A1[0], A0[1]);
#pragma unroll
  for (int w = 1; w < N - 1; w++) r.v[w] = addc_cc(A1[w], A0[w + 1]);
  r.v[N - 1] = addc(A1[N - 1], A0[N]);
  fe_reduce_once(r);
  return r;
}

// Out-of-line product for the latency-bound tail kernels (one warp or a few): their bodies are
// dozens of inlined products otherwise (70-220 KB of SASS against a 32 KB instruction cache).
#ifdef __CUDACC__
template <class F>
__device__ __noinline__ Fe<F> fe_mul_call(Fe<F> a, Fe<F> b) {
  return fe_mul(a, b);
}
#else
template <class F>
inline Fe<F> fe_mul_call(const Fe<F>& a, const Fe<F>& b) {
  return fe_mul(a, b);
}
#endif

// One Montgomery reduction row on the split accumulator: mont_row with the a*b part removed.
template <class F, bool FIRST>
MSM_HD void redc_row(uint32_t* U, uint32_t* V) {
  constexpr int N = F::N;
  if (!FIRST) {
    V[0] = add_cc(V[0], U[1]);
#pragma unroll
    for (int j = 0; j < N - 2; j++) U[j] = addc_cc(U[j + 2], 0u);
    U[N - 2] = addc_cc(U[N], 0u);
    U[N - 1] = addc(0u, 0u);
  }
  V[N] = 0;
  uint32_t m = mul_lo(V[0], F::M0v());
  if (FIRST) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
      mul_wide(F::P(j + 1), m, U[j], U[j + 1]);
    }
  } else {
    U[0] = mad_lo_cc(F::P(1), m, U[0]);
    U[1] = madc_hi_cc(F::P(1), m, U[1]);
#pragma unroll
    for (int j = 2; j < N - 2; j += 2) {
      U[j] = madc_lo_cc(F::P(j + 1), m, U[j]);
      U[j + 1] = madc_hi_cc(F::P(j + 1), m, U[j + 1]);
    }
    if (N > 2) {
      U[N - 2] = madc_lo_cc(F::P(N - 1), m, U[N - 2]);
      U[N - 1] = madc_hi(F::P(N - 1), m, U[N - 1]);
    }
  }
  V[0] = mad_lo_cc(F::P(0), m, V[0]);
  V[1] = madc_hi_cc(F::P(0), m, V[1]);
#pragma unroll
  for (int j = 2; j < N; j += 2) {
    V[j] = madc_lo_cc(F::P(j), m, V[j]);
    V[j + 1] = madc_hi_cc(F::P(j), m, V[j + 1]);
  }
  V[N] = addc(V[N], 0u);
}

// a^2 / R.  The N(N-1)/2 cross products a_i*a_j (i<j) are accumulated once into two word-aligned
// planes (E: i+j even, O: i+j odd, held one word up) so that every lo/hi pair still fuses into one
// wide multiply-add; the sum is doubled, the N squares added, and the 2N-word result reduced with
// N redc rows: N(N+1)/2 + N^2 wide products against 2N^2 for fe_mul.
// Carry-outs of the per-row chains land in a word that so far holds at most another chain's carry
// (row i ends one pair beyond row i-1 in each plane), so a plain addc is exact.
template <class F>
MSM_HD Fe<F> fe_sqr(const Fe<F>& a) {
  constexpr int N = F::N;
  static_assert(N % 2 == 0, "even limb count");
  uint32_t E[2 * N], O[2 * N];
#pragma unroll
  for (int k = 0; k < 2 * N; k++) E[k] = O[k] = 0;
#pragma unroll
  for (int i = 0; i < N - 1; i++) {
    {  // j - i odd: word i+j is odd, O index i+j-1
      int last = 0;
#pragma unroll
      for (int j = i + 1; j < N; j += 2) {
        int k = i + j - 1;
        O[k] = (j == i + 1) ? mad_lo_cc(a.v[i], a.v[j], O[k]) : madc_lo_cc(a.v[i], a.v[j], O[k]);
        O[k + 1] = madc_hi_cc(a.v[i], a.v[j], O[k + 1]);
        last = k + 1;
      }
      O[last + 1] = addc(O[last + 1], 0u);
    }
    if (i + 2 < N) {  // j - i even: E index i+j
      int last = 0;
#pragma unroll
      for (int j = i + 2; j < N; j += 2) {
        int k = i + j;
        E[k] = (j == i + 2) ? mad_lo_cc(a.v[i], a.v[j], E[k]) : madc_lo_cc(a.v[i], a.v[j], E[k]);
        E[k + 1] = madc_hi_cc(a.v[i], a.v[j], E[k + 1]);
        last = k + 1;
      }
      E[last + 1] = addc(E[last + 1], 0u);
    }
  }
  // C = E + (O one word up); word 0 is empty
  uint32_t D[2 * N];
  D[0] = 0;
  D[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 2 * N - 1; k++) D[k] = addc_cc(E[k], O[k - 1]);
  D[2 * N - 1] = addc(E[2 * N - 1], O[2 * N - 2]);
  // doubled, plus the squares
#pragma unroll
  for (int k = 2 * N - 1; k >= 1; k--) D[k] = (D[k] << 1) | (D[k - 1] >> 31);
  D[0] = mad_lo_cc(a.v[0], a.v[0], D[0]);
  D[1] = madc_hi_cc(a.v[0], a.v[0], D[1]);
#pragma unroll
  for (int i = 1; i < N - 1; i++) {
    D[2 * i] = madc_lo_cc(a.v[i], a.v[i], D[2 * i]);
    D[2 * i + 1] = madc_hi_cc(a.v[i], a.v[i], D[2 * i + 1]);
  }
  D[2 * N - 2] = madc_lo_cc(a.v[N - 1], a.v[N - 1], D[2 * N - 2]);
  D[2 * N - 1] = madc_hi(a.v[N - 1], a.v[N - 1], D[2 * N - 1]);
  // reduce the low half, N rows
  uint32_t A0[N + 1], A1[N + 1];
  A0[N] = 0;
#pragma unroll
  for (int k = 0; k < N; k++) A1[k] = D[k];
  redc_row<F, true>(A0, A1);  // X = A1, Y = A0
#pragma unroll
  for (int i = 1; i < N; i += 2) {
    redc_row<F, false>(A1, A0);
    if (i + 1 < N) redc_row<F, false>(A0, A1);
  }
  // X = A0, Y = A1:  (Y + (X >> 32)) <= p, plus the high half (< p)
  Fe<F> r;
  r.v[0] = add_cc(A1[0], A0[1]);
#pragma unroll
  for (int w = 1; w < N - 1; w++) r.v[w] = addc_cc(A1[w], A0[w + 1]);
  r.v[N - 1] = addc(A1[N - 1], A0[N]);
  r.v[0] = add_cc(r.v[0], D[N]);
#pragma unroll
  for (int w = 1; w < N - 1; w++) r.v[w] = addc_cc(r.v[w], D[N + w]);
  r.v[N - 1] = addc(r.v[N - 1], D[2 * N - 1]);
  fe_reduce_once(r);
  return r;
}

#ifdef __CUDACC__
template <class F>
__device__ __noinline__ Fe<F> fe_sqr_call(Fe<F> a) {
  return fe_sqr(a);
}
#else
template <class F>
inline Fe<F> fe_sqr_call(const Fe<F>& a) {
  return fe_sqr(a);
}
#endif

// ------------------------------------------------------------------------------------------
// Unreduced ("lazy") helpers for fields with F::LAZY (floor(2^(32N) / p) >= 64, i.e. BLS12-377 Fq with its 7
// spare bits).  Values are bounded multiples of p that still fit the limbs; no conditional subtraction, and
// the shifts have no carry chain at all.  fe_mul / fe_sqr accept such operands as long as
// (a/p) * (b/p) < floor(2^(32N) / p) and return a canonical value.  Used by the one-warp tail only
// (kernels_reduce.cuh), where every dependent instruction costs ~4.5 cycles.
// ------------------------------------------------------------------------------------------
template <class F>
MSM_HD Fe<F> fe_add_nr(const Fe<F>& a, const Fe<F>& b) {  // a + b, caller guarantees a + b < 2^(32N)
  Fe<F> r;
  r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < F::N - 1; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
  r.v[F::N - 1] = addc(a.v[F::N - 1], b.v[F::N - 1]);
  return r;
}

template <class F, int S>
MSM_HD Fe<F> fe_shl_nr(const Fe<F>& a) {  // a * 2^S, S in 1..3, no overflow by the caller's bound
  Fe<F> r;
#pragma unroll
  for (int i = F::N - 1; i >= 1; i--) r.v[i] = (a.v[i] << S) | (a.v[i - 1] >> (32 - S));
  r.v[0] = a.v[0] << S;
  return r;
}

// a - b + K p for K in {1, 2, 3, 9}: positive whenever b < K p
template <class F, int K>
MSM_HD Fe<F> fe_sub_nr(const Fe<F>& a, const Fe<F>& b) {
  Fe<F> t;
  t.v[0] = sub_cc(K == 1 ? F::P(0) : (K == 2 ? F::P2(0) : (K == 3 ? F::P3(0) : F::P9(0))), b.v[0]);
#pragma unroll
  for (int i = 1; i < F::N - 1; i++)
    t.v[i] = subc_cc(K == 1 ? F::P(i) : (K == 2 ? F::P2(i) : (K == 3 ? F::P3(i) : F::P9(i))), b.v[i]);
  t.v[F::N - 1] = subc(K == 1 ? F::P(F::N - 1) : (K == 2 ? F::P2(F::N - 1) : (K == 3 ? F::P3(F::N - 1) : F::P9(F::N - 1))),
                       b.v[F::N - 1]);
  return fe_add_nr(a, t);
}

// a * (small unsigned constant), by double-and-add on the constant's bits (c >= 1)
template <class F>
MSM_HD Fe<F> fe_mul_small(const Fe<F>& a, uint32_t c) {
  Fe<F> r = a;
  int top = 31;
  while (!((c >> top) & 1)) top--;
  for (int i = top - 1; i >= 0; i--) {
    r = fe_dbl(r);
    if ((c >> i) & 1) r = fe_add(r, a);
  }
  return r;
}

// Montgomery-domain inverse by Fermat: a^(p-2)   (a = x*R  ->  x^-1 * R).  a != 0.
// Fixed 4-bit windows over the exponent p-2.  Kept as the independent cross-check of fe_inv
// (tests/test_hostmath.py); ~480 dependent Montgomery products = ~0.3 ms for a lone warp.
template <class F>
MSM_HD Fe<F> fe_inv_fermat(const Fe<F>& a) {
  Fe<F> tbl[16];
  tbl[0] = fe_one<F>();
  tbl[1] = a;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tbl[i] = fe_mul(tbl[i - 1], a);
  Fe<F> r = fe_one<F>();
  bool started = false;
#pragma unroll 1
  for (int i = F::N * 8 - 1; i >= 0; i--) {
    uint32_t nib = (F::PM2(i >> 3) >> ((i & 7) * 4)) & 15u;
    if (started) {
      r = fe_sqr(r);
      r = fe_sqr(r);
      r = fe_sqr(r);
      r = fe_sqr(r);
    }
    if (nib) {
      r = started ? fe_mul(r, tbl[nib]) : tbl[nib];
      started = true;
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------------
// Montgomery-domain inverse, a = x*R -> x^-1 * R  (a != 0): variable-time "safegcd" divsteps
// (Bernstein-Yang 2019) in batches of 30 on signed 30-bit limbs, followed by one Montgomery
// product with R^3.  Replaces the reference's Kaliski almost-inverse + fix-up
// (src/wasm/inverse.ts:42-218).  ~27 batches of (30 divsteps on the low words + two linear
// updates of L30 limbs) instead of ~480 dependent field multiplications: this is what sits on the
// critical path of every batched-affine round (one true inversion per product tree).
// ------------------------------------------------------------------------------------------
MSM_HD int msm_ctz32(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

template <class F>
struct Inv30 {
  static constexpr int L = (F::BITS + 2 + 29) / 30;  // limbs: values range over (-2p, p)
  static constexpr int32_t M30 = 0x3FFFFFFF;
  int32_t v[L];
};

template <class F>
MSM_HD int32_t inv30_modulus_limb(int i) {  // bits [30i, 30i+30) of p
  const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
  uint32_t lo = (k < F::N) ? (F::P(k) >> sh) : 0u;
  if (sh > 2 && k + 1 < F::N) lo |= F::P(k + 1) << (32 - sh);
  return (int32_t)(lo & 0x3FFFFFFFu);
}

template <class F>
MSM_HD Fe<F> fe_inv_plain(const Fe<F>& x) {  // x^-1 mod p as plain integers, x in [1, p)
  typedef Inv30<F> I;
  constexpr int L = I::L;
  constexpr int32_t M30 = I::M30;
  int32_t P30[L];
#pragma unroll
  for (int i = 0; i < L; i++) P30[i] = inv30_modulus_limb<F>(i);
  // p^-1 mod 2^30 by Newton iteration on the low word
  uint32_t pinv = P30[0];
#pragma unroll
  for (int i = 0; i < 5; i++) pinv *= 2u - (uint32_t)P30[0] * pinv;
  pinv &= (uint32_t)M30;

  int32_t d[L], e[L], f[L], g[L];
#pragma unroll
  for (int i = 0; i < L; i++) {
    d[i] = 0;
    e[i] = 0;
    f[i] = P30[i];
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    uint32_t lo = (k < F::N) ? (x.v[k] >> sh) : 0u;
    if (sh > 2 && k + 1 < F::N) lo |= x.v[k + 1] << (32 - sh);
    g[i] = (int32_t)(lo & (uint32_t)M30);
  }
  e[0] = 1;
  int32_t eta = -1;
#pragma unroll 1
  for (int iter = 0; iter < 64; iter++) {
    // ---- 30 divsteps on the low 32 bits of f, g -> transition matrix (u v; q r)
    uint32_t u = 1, v = 0, q = 0, r = 1;
    uint32_t f0 = (uint32_t)f[0] | ((uint32_t)f[1] << 30), g0 = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
    int i = 30;
#pragma unroll 1
    for (;;) {
      int zeros = msm_ctz32(g0 | (0xFFFFFFFFu << i));
      g0 >>= zeros;
      u <<= zeros;
      v <<= zeros;
      eta -= zeros;
      i -= zeros;
      if (i == 0) break;
      if (eta < 0) {
        uint32_t t;
        eta = -eta;
        t = f0, f0 = g0, g0 = 0u - t;
        t = u, u = q, q = 0u - t;
        t = v, v = r, r = 0u - t;
      }
      int limit = (eta + 1) > i ? i : (eta + 1);
      uint32_t m = (0xFFFFFFFFu >> (32 - limit)) & 63u;
      uint32_t w = (f0 * g0 * (f0 * f0 - 2u)) & m;  // -g/f mod 2^min(limit,6)
      g0 += f0 * w;
      q += u * w;
      r += v * w;
    }
    const int32_t su = (int32_t)u, sv = (int32_t)v, sq = (int32_t)q, sr = (int32_t)r;
    // ---- (d, e) <- (u v; q r) (d, e) / 2^30 mod p
    {
      int32_t sd = d[L - 1] >> 31, se = e[L - 1] >> 31;
      int32_t md = (su & sd) + (sv & se), me = (sq & sd) + (sr & se);
      int64_t cd = (int64_t)su * d[0] + (int64_t)sv * e[0];
      int64_t ce = (int64_t)sq * d[0] + (int64_t)sr * e[0];
      md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      me -= (int32_t)((pinv * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
      cd += (int64_t)P30[0] * md;
      ce += (int64_t)P30[0] * me;
      cd >>= 30;
      ce >>= 30;
#pragma unroll
      for (int k = 1; k < L; k++) {
        cd += (int64_t)su * d[k] + (int64_t)sv * e[k];
        ce += (int64_t)sq * d[k] + (int64_t)sr * e[k];
        cd += (int64_t)P30[k] * md;
        ce += (int64_t)P30[k] * me;
        d[k - 1] = (int32_t)cd & M30;
        cd >>= 30;
        e[k - 1] = (int32_t)ce & M30;
        ce >>= 30;
      }
      d[L - 1] = (int32_t)cd;
      e[L - 1] = (int32_t)ce;
    }
    // ---- (f, g) <- (u v; q r) (f, g) / 2^30
    {
      int64_t cf = (int64_t)su * f[0] + (int64_t)sv * g[0];
      int64_t cg = (int64_t)sq * f[0] + (int64_t)sr * g[0];
      cf >>= 30;
      cg >>= 30;
#pragma unroll
      for (int k = 1; k < L; k++) {
        cf += (int64_t)su * f[k] + (int64_t)sv * g[k];
        cg += (int64_t)sq * f[k] + (int64_t)sr * g[k];
        f[k - 1] = (int32_t)cf & M30;
        cf >>= 30;
        g[k - 1] = (int32_t)cg & M30;
        cg >>= 30;
      }
      f[L - 1] = (int32_t)cf;
      g[L - 1] = (int32_t)cg;
    }
    int32_t nz = 0;
#pragma unroll
    for (int k = 0; k < L; k++) nz |= g[k];
    if (nz == 0) break;
  }
  // f = +-1 now; result = sign(f) * d, normalised to [0, p)
  {
    int32_t sign = f[L - 1];
    int32_t cond_add = d[L - 1] >> 31;
    int32_t cond_neg = sign >> 31;
#pragma unroll
    for (int k = 0; k < L; k++) {
      d[k] += P30[k] & cond_add;
      d[k] = (d[k] ^ cond_neg) - cond_neg;
    }
#pragma unroll
    for (int k = 0; k < L - 1; k++) {
      d[k + 1] += d[k] >> 30;
      d[k] &= M30;
    }
    cond_add = d[L - 1] >> 31;
#pragma unroll
    for (int k = 0; k < L; k++) d[k] += P30[k] & cond_add;
#pragma unroll
    for (int k = 0; k < L - 1; k++) {
      d[k + 1] += d[k] >> 30;
      d[k] &= M30;
    }
  }
  Fe<F> out;
#pragma unroll
  for (int k = 0; k < F::N; k++) out.v[k] = 0;
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    uint32_t l = (uint32_t)d[i];
    if (k < F::N) out.v[k] |= l << sh;
    if (sh > 2 && k + 1 < F::N) out.v[k + 1] |= l >> (32 - sh);
  }
  return out;
}

template <class F>
MSM_HD Fe<F> fe_inv(const Fe<F>& a) {
  Fe<F> r3;
#pragma unroll
  for (int i = 0; i < F::N; i++) r3.v[i] = F::R3(i);
  return fe_mul(fe_inv_plain(a), r3);  // (x R)^-1 * R^3 / R = x^-1 R
}

// x -> x*R (to Montgomery) and back
template <class F>
MSM_HD Fe<F> fe_to_mont(const Fe<F>& a) {
  Fe<F> r2;
#pragma unroll
  for (int i = 0; i < F::N; i++) r2.v[i] = F::R2(i);
  return fe_mul(a, r2);
}

template <class F>
MSM_HD Fe<F> fe_from_mont(const Fe<F>& a) {
  Fe<F> one = fe_zero<F>();
  one.v[0] = 1;
  return fe_mul(a, one);
}

}  // namespace msm
