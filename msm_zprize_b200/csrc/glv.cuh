// Scalar side: GLV decomposition and signed-window digits on 32-bit limbs.
//
// Replaces the reference's scalar wasm module:
//   decompose         src/wasm/glv.ts:68-169 (constants :45-63, multiplyMsb rounding :187-214),
//                     BigInt restatement src/glv/glv-test.ts:96-100,143-149
//   extractBitSlice   src/wasm/field-helpers.ts:307-358
//   digit rule        src/msm-batched-affine.ts:178-191 / src/msm-basic.ts:83-92
// The decomposition reproduces the reference's (same m, k, lattice basis, rounding), so the
// half-scalars and digits can be compared one to one with the oracle.
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace msm {

// z[0..NZ) = low NZ limbs of x[0..NX) * y[0..NY)
template <int NX, int NY, int NZ>
MSM_HD void mp_mul_trunc(uint32_t* z, const uint32_t* x, const uint32_t* y) {
  uint64_t acc_lo = 0;  // running column sum, 96-bit as (acc_hi : acc_lo)
  uint32_t acc_hi = 0;
#pragma unroll
  for (int k = 0; k < NZ; k++) {
#pragma unroll
    for (int i = 0; i < NX; i++) {
      int j = k - i;
      if (j < 0 || j >= NY) continue;
      uint64_t pr = (uint64_t)x[i] * y[j];
      uint64_t s = acc_lo + pr;
      acc_hi += (s < pr) ? 1u : 0u;
      acc_lo = s;
    }
    z[k] = (uint32_t)acc_lo;
    acc_lo = (acc_lo >> 32) | ((uint64_t)acc_hi << 32);
    acc_hi = 0;
  }
}

template <int N>
MSM_HD void mp_add(uint32_t* z, const uint32_t* x) {  // z += x mod 2^(32N)
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    c += (uint64_t)z[i] + x[i];
    z[i] = (uint32_t)c;
    c >>= 32;
  }
}

template <int N>
MSM_HD void mp_sub(uint32_t* z, const uint32_t* x) {  // z -= x mod 2^(32N)
  uint64_t b = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    uint64_t t = (uint64_t)z[i] - x[i] - b;
    z[i] = (uint32_t)t;
    b = (t >> 32) & 1;
  }
}

template <int N>
MSM_HD void mp_neg(uint32_t* z) {  // two's complement
  uint64_t c = 1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    c += (uint64_t)(~z[i]);
    z[i] = (uint32_t)c;
    c >>= 32;
  }
}

// s >= q ?  (8 limbs)
template <class S>
MSM_HD bool scalar_geq_q(const uint32_t* s) {
#pragma unroll
  for (int i = 7; i >= 0; i--) {
    uint32_t qi = S::Q(i);
    if (s[i] > qi) return true;
    if (s[i] < qi) return false;
  }
  return true;
}

// bring an arbitrary 256-bit value into [0, q) by repeated subtraction (inputs are specified
// < q, src/scripts/msm-weierstrass.ts:74-78; this keeps out-of-range inputs well defined)
template <class S>
MSM_HD void scalar_reduce(uint32_t* s) {
  for (int it = 0; it < 64 && scalar_geq_q<S>(s); it++) {
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = S::Q(i);
    mp_sub<8>(s, q);
  }
}

// x = round(|m| * (s >> k) / 2^m)   with |m| 5 limbs, result 5 limbs
template <class G>
MSM_HD void glv_round_mul(uint32_t* x, const uint32_t* sh, const uint32_t* mabs) {
  uint32_t prod[10];
  mp_mul_trunc<5, 5, 10>(prod, sh, mabs);
  // bits [M, M+160) plus rounding bit M-1
  constexpr int M = G::SHIFT_M;
  constexpr int w = M / 32, b = M % 32;
  static_assert(b != 0, "shift assumed not word aligned");
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint32_t lo = (w + i < 10) ? prod[w + i] : 0u;
    uint32_t hi = (w + i + 1 < 10) ? prod[w + i + 1] : 0u;
    x[i] = (lo >> b) | (hi << (32 - b));
  }
  uint32_t rnd = (prod[(M - 1) / 32] >> ((M - 1) % 32)) & 1u;
  uint64_t c = rnd;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    c += x[i];
    x[i] = (uint32_t)c;
    c >>= 32;
  }
}

// acc (5 limbs, mod 2^160) +=/-= |v| * x
template <int NEG>
MSM_HD void glv_acc_term(uint32_t* acc, const uint32_t* vabs, const uint32_t* x) {
  uint32_t t[5];
  mp_mul_trunc<5, 5, 5>(t, vabs, x);
  if (NEG)
    mp_sub<5>(acc, t);
  else
    mp_add<5>(acc, t);
}

// s (8 limbs, < q)  ->  |s0|, |s1| (4 limbs each, < 2^127) and sign flags (bit0: s0 < 0, bit1: s1 < 0)
template <class G>
MSM_HD uint32_t glv_decompose(const uint32_t* s, uint32_t* s0, uint32_t* s1) {
  constexpr int K = G::SHIFT_K;
  constexpr int kw = K / 32, kb = K % 32;
  static_assert(kb != 0, "shift assumed not word aligned");
  uint32_t sh[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint32_t lo = (kw + i < 8) ? s[kw + i] : 0u;
    uint32_t hi = (kw + i + 1 < 8) ? s[kw + i + 1] : 0u;
    sh[i] = (lo >> kb) | (hi << (32 - kb));
  }
  uint32_t m0[5], m1[5], v00[5], v01[5], v10[5], v11[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    m0[i] = G::M0(i);
    m1[i] = G::M1(i);
    v00[i] = G::V00(i);
    v01[i] = G::V01(i);
    v10[i] = G::V10(i);
    v11[i] = G::V11(i);
  }
  uint32_t x0[5], x1[5];
  glv_round_mul<G>(x0, sh, m0);  // |x0|, sign(x0) = sign(m0)
  glv_round_mul<G>(x1, sh, m1);
  // s0 = v00*x0 + v01*x1 + s ; s1 = v10*x0 + v11*x1      (mod 2^160, then sign from bit 159)
  uint32_t a0[5], a1[5];
#pragma unroll
  for (int i = 0; i < 5; i++) {
    a0[i] = s[i];
    a1[i] = 0;
  }
  glv_acc_term<G::V00_NEG ^ G::M0_NEG>(a0, v00, x0);
  glv_acc_term<G::V01_NEG ^ G::M1_NEG>(a0, v01, x1);
  glv_acc_term<G::V10_NEG ^ G::M0_NEG>(a1, v10, x0);
  glv_acc_term<G::V11_NEG ^ G::M1_NEG>(a1, v11, x1);
  uint32_t flags = 0;
  if (a0[4] >> 31) {
    mp_neg<5>(a0);
    flags |= 1u;
  }
  if (a1[4] >> 31) {
    mp_neg<5>(a1);
    flags |= 2u;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    s0[i] = a0[i];
    s1[i] = a1[i];
  }
  return flags;
}

// `len` (<= 31) bits of x[0..NW) starting at bit `start`; bits beyond the top limb read as zero
template <int NW>
MSM_HD uint32_t extract_bits(const uint32_t* x, int start, int len) {
  int w = start >> 5, b = start & 31;
  uint32_t lo = (w < NW) ? x[w] : 0u;
  uint32_t hi = (w + 1 < NW) ? x[w + 1] : 0u;
  uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> b) & ((1u << len) - 1u);
}

}  // namespace msm
