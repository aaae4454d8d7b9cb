// Device-side plumbing shared by the MSM kernels: memory layouts, loaders, digit extraction.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "constants.cuh"
#include "ec.cuh"
#include "glv.cuh"

namespace msm {

// ------------------------------------------------------------------------------------------
// Layouts in HBM
//
//  * "SoA chunk" arrays of field elements: element i, 16-byte chunk c (4 limbs) lives at
//    base[c * stride + i] (uint4).  Consecutive threads touch consecutive 16-byte words: every
//    load/store of the batched-affine kernels is a fully coalesced 512-byte warp transaction.
//  * base points: AoS records of 2N limbs (x | y), 96 bytes (N=12) or 64 bytes (N=8): a random
//    gather reads whole 32-byte sectors.  Record 2i is G_i, record 2i+1 is endo(G_i) = (beta x, y)
//    (the reference stores 4 variants incl. negations, src/msm-batched-affine.ts:338-409; the
//    negation is applied on the fly here).
// ------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ Fe<F> ld_soa(const uint4* __restrict__ base, size_t stride, size_t i) {
  Fe<F> r;
#pragma unroll
  for (int c = 0; c < F::N / 4; c++) {
    uint4 q = base[c * stride + i];
    r.v[4 * c + 0] = q.x;
    r.v[4 * c + 1] = q.y;
    r.v[4 * c + 2] = q.z;
    r.v[4 * c + 3] = q.w;
  }
  return r;
}

template <class F>
__device__ __forceinline__ void st_soa(uint4* __restrict__ base, size_t stride, size_t i, const Fe<F>& a) {
#pragma unroll
  for (int c = 0; c < F::N / 4; c++)
    base[c * stride + i] = make_uint4(a.v[4 * c], a.v[4 * c + 1], a.v[4 * c + 2], a.v[4 * c + 3]);
}

template <class F>
__device__ __forceinline__ Fe<F> ld_aos(const uint4* __restrict__ p) {
  Fe<F> r;
#pragma unroll
  for (int c = 0; c < F::N / 4; c++) {
    uint4 q = __ldg(p + c);
    r.v[4 * c + 0] = q.x;
    r.v[4 * c + 1] = q.y;
    r.v[4 * c + 2] = q.z;
    r.v[4 * c + 3] = q.w;
  }
  return r;
}

template <class F>
__device__ __forceinline__ void st_aos(uint4* p, const Fe<F>& a) {
#pragma unroll
  for (int c = 0; c < F::N / 4; c++)
    p[c] = make_uint4(a.v[4 * c], a.v[4 * c + 1], a.v[4 * c + 2], a.v[4 * c + 3]);
}

// Elements of one accumulation round: slot e -> plane (e & 1), index (e >> 1); each plane holds an
// x and a y SoA-chunk array of `cap` elements.  Pair i = slots (2i, 2i+1) = index i of both planes.
template <class F>
struct ElemBuf {
  uint4* base;
  size_t cap;
  static constexpr int CH = F::N / 4;
  __device__ __forceinline__ uint4* coord(int plane, int xy) const {
    return base + (size_t)((plane * 2 + xy) * CH) * cap;
  }
  __device__ __forceinline__ Aff<F> load(size_t e) const {
    Aff<F> P;
    P.x = ld_soa<F>(coord((int)(e & 1), 0), cap, e >> 1);
    P.y = ld_soa<F>(coord((int)(e & 1), 1), cap, e >> 1);
    return P;
  }
  __device__ __forceinline__ void store(size_t e, const Aff<F>& P) const {
    st_soa<F>(coord((int)(e & 1), 0), cap, e >> 1, P.x);
    st_soa<F>(coord((int)(e & 1), 1), cap, e >> 1, P.y);
  }
  static size_t bytes(size_t cap) { return (size_t)4 * CH * cap * sizeof(uint4); }
};

// entry of the sorted index array: half-scalar / point index in bits 0..30, negate flag in bit 31
__device__ __forceinline__ uint32_t ent_index(uint32_t e) { return e & 0x7FFFFFFFu; }
__device__ __forceinline__ bool ent_neg(uint32_t e) { return (e >> 31) != 0; }

template <class F>
__device__ __forceinline__ Aff<F> gather_base(const uint4* __restrict__ bases, uint32_t e) {
  const uint4* p = bases + (size_t)ent_index(e) * (2 * F::N / 4);
  Aff<F> P;
  P.x = ld_aos<F>(p);
  P.y = ld_aos<F>(p + F::N / 4);
  if (ent_neg(e) && !aff_is_inf(P)) P.y = fe_neg(P.y);
  return P;
}

// Signed-window digit k of an NW-limb magnitude (src/msm-batched-affine.ts:178-191):
// returns l in [0, L]; carry in/out says "the point enters bucket l negated".
template <int NW>
__device__ __forceinline__ uint32_t signed_digit(const uint32_t* s, int k, int c, uint32_t& carry) {
  uint32_t L = 1u << (c - 1);
  uint32_t l = extract_bits<NW>(s, k * c, c) + carry;
  if (l > L) {
    l = 2 * L - l;
    carry = 1;
  } else {
    carry = 0;
  }
  return l;
}

}  // namespace msm
