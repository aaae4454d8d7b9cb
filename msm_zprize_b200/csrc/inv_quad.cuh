// Quad-cooperative safegcd inversion for the serial spots of the pipeline (the top of the product tree
// in every batched-affine round, the final normalisation): the same algorithm as fe_inv_plain (fp.cuh:
// variable-time Bernstein-Yang divsteps in batches of 30 on signed 30-bit limbs), with the four linear
// updates of every batch spread over the four lanes of a quad.
//
//   lane q = 0: owns d      1: owns e      2: owns f      3: owns g
//   every lane keeps its own vector A and a copy B of its partner's (lane q ^ 1), so that
//       A' = (c0 * A + c1 * B [+ m * p]) / 2^30     with (c0, c1) = (u, v) for d and f, (r, q) for e and g
//   is one chain per lane; afterwards the partners swap their new vectors (L shuffles), lanes 2 / 3
//   broadcast the low words of f / g (4 shuffles) and everyone runs the next 30 divsteps on them.
// A lone thread spends ~560 dependent instructions per batch, a quad lane ~270 (measured: 43 -> see
// DESIGN.md).  All 32 lanes of the calling warp must be converged; every quad computes the same
// thing, the result is valid in lane 0 of each quad (and returned in every lane).
#pragma once
#include "fp.cuh"

namespace msm {

template <class F>
__device__ __forceinline__ Fe<F> fe_inv_plain_quad(const Fe<F>& x) {  // x^-1 mod p, plain integers, x in [1, p)
  typedef Inv30<F> I;
  constexpr int L = I::L;
  constexpr int32_t M30 = I::M30;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, q = lane & 3, base = lane & ~3;
  const bool is_fg = q >= 2;   // lanes 2, 3 carry f, g (no modular correction)
  const bool second = q & 1;   // e or g: coefficients (r, q) instead of (u, v)
  int32_t P30[L];
#pragma unroll
  for (int i = 0; i < L; i++) P30[i] = inv30_modulus_limb<F>(i);
  uint32_t pinv = P30[0];
#pragma unroll
  for (int i = 0; i < 5; i++) pinv *= 2u - (uint32_t)P30[0] * pinv;
  pinv &= (uint32_t)M30;

  // A = own vector, B = partner's:  d = 0, e = 1, f = p, g = x
  int32_t A[L], B[L];
#pragma unroll
  for (int i = 0; i < L; i++) {
    const int bit = 30 * i, k = bit >> 5, sh = bit & 31;
    uint32_t lo = (k < F::N) ? (x.v[k] >> sh) : 0u;
    if (sh > 2 && k + 1 < F::N) lo |= x.v[k + 1] << (32 - sh);
    const int32_t gx = (int32_t)(lo & (uint32_t)M30);
    const int32_t one = (i == 0) ? 1 : 0;
    // q: 0 -> (d, e) = (0, 1)   1 -> (e, d) = (1, 0)   2 -> (f, g) = (p, x)   3 -> (g, f) = (x, p)
    A[i] = q == 0 ? 0 : (q == 1 ? one : (q == 2 ? P30[i] : gx));
    B[i] = q == 0 ? one : (q == 1 ? 0 : (q == 2 ? gx : P30[i]));
  }
  int32_t eta = -1;
#pragma unroll 1
  for (int iter = 0; iter < 64; iter++) {
    // low words of f and g for everyone
    const uint32_t a01 = (uint32_t)A[0] | ((uint32_t)A[1] << 30);
    uint32_t f0 = __shfl_sync(FULL, a01, base + 2), g0 = __shfl_sync(FULL, a01, base + 3);
    uint32_t u = 1, v = 0, qq = 0, r = 1;
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
        t = u, u = qq, qq = 0u - t;
        t = v, v = r, r = 0u - t;
      }
      int limit = (eta + 1) > i ? i : (eta + 1);
      uint32_t m = (0xFFFFFFFFu >> (32 - limit)) & 63u;
      uint32_t w = (f0 * g0 * (f0 * f0 - 2u)) & m;
      g0 += f0 * w;
      qq += u * w;
      r += v * w;
    }
    // this lane's row of the transition matrix, applied to (A, B):
    //   d' = u d + v e    e' = q d + r e = r e + q d    f' = u f + v g    g' = q f + r g = r g + q f
    const int32_t c0 = second ? (int32_t)r : (int32_t)u;
    const int32_t c1 = second ? (int32_t)qq : (int32_t)v;
    {
      int32_t sA = A[L - 1] >> 31, sB = B[L - 1] >> 31;
      int32_t md = (c0 & sA) + (c1 & sB);
      int64_t cd = (int64_t)c0 * A[0] + (int64_t)c1 * B[0];
      md -= (int32_t)((pinv * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
      if (is_fg) md = 0;  // f, g: exact division by 2^30, no multiple of p
      cd += (int64_t)P30[0] * md;
      cd >>= 30;
#pragma unroll
      for (int k = 1; k < L; k++) {
        cd += (int64_t)c0 * A[k] + (int64_t)c1 * B[k];
        cd += (int64_t)P30[k] * md;
        A[k - 1] = (int32_t)cd & M30;
        cd >>= 30;
      }
      A[L - 1] = (int32_t)cd;
    }
    // partners swap their new vectors
#pragma unroll
    for (int k = 0; k < L; k++) B[k] = __shfl_xor_sync(FULL, A[k], 1);
    // g == 0 ?  (lane 3 owns g)
    int32_t nz = 0;
#pragma unroll
    for (int k = 0; k < L; k++) nz |= A[k];
    nz = __shfl_sync(FULL, nz, base + 3);
    if (nz == 0) break;
  }
  // f = +-1 now; result = sign(f) * d, normalised to [0, p): computed from lane 0's d and lane 2's sign of f
  const int32_t fsign = __shfl_sync(FULL, A[L - 1], base + 2);
  int32_t d[L];
#pragma unroll
  for (int k = 0; k < L; k++) d[k] = __shfl_sync(FULL, A[k], base);
  {
    int32_t cond_add = d[L - 1] >> 31;
    int32_t cond_neg = fsign >> 31;
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

// Montgomery-domain inverse, all 32 lanes of the warp with the same argument
template <class F>
__device__ __forceinline__ Fe<F> fe_inv_quad(const Fe<F>& a) {
  Fe<F> r3;
#pragma unroll
  for (int i = 0; i < F::N; i++) r3.v[i] = F::R3(i);
  return fe_mul(fe_inv_plain_quad(a), r3);
}

}  // namespace msm
