// Kernels of the GLV + batched-affine Pippenger MSM on short Weierstrass curves (a = 0).
// GPU redesign of src/msm-batched-affine.ts:74-328 (reference paths relative to its repo):
//
//   k_ingest_points   preparePointsAndScalars (points half), :338-409 + Parallel.pointsFromBytes
//   k_glv             decompose, src/wasm/glv.ts:68-169
//   k_hist            "slice scalars & count buckets", :166-202
//   k_scan            integrateBucketCounts, :411-435 (for every tree round at once)
//   k_scatter         sortPoints, :444-490 -- sorts 32-bit indices, not 116-byte points
//   k_fwd / k_bwd     the accumulation loop :226-271 with batchAddNew, src/curve-affine.ts:376-458:
//                     one pairwise-tree round = forward product pass, batched inversion, backward
//                     pass that finishes the affine additions
//   k_up_* / k_inv    the batched inversion itself (src/wasm/inverse.ts:220-271 `batchInverse`),
//                     as a multi-level product tree with one Fermat inversion per top element
#pragma once
#include "kernels_common.cuh"

namespace msm {

constexpr int ACC_THREADS = 256;  // block size of the level-0 kernels
constexpr int ACC_B0 = 8;         // pairs per thread, level 0
constexpr int UP_THREADS = 32;    // block size of the upper product-tree levels
constexpr int UP_B1 = 32;         // elements per thread, upper levels
constexpr int TOP_MAX = 2048;     // at most this many Fermat inversions per round
constexpr int MAX_ROUNDS = 30;

// ------------------------------------------------------------------------------------------
// ingest
// ------------------------------------------------------------------------------------------

// 29-bit limbs in u32 words (Montgomery R29) -> canonical Montgomery R32.  The input may be
// unreduced in [0, 2p) (wasm `multiply` only guarantees < 2p).
template <class F>
__device__ __forceinline__ Fe<F> fe_from_limb29(const uint32_t* __restrict__ w) {
  Fe<F> r = fe_zero<F>();
#pragma unroll
  for (int j = 0; j < F::N29; j++) {
    uint32_t l = w[j] & 0x1FFFFFFFu;
    int bit = 29 * j;
    int word = bit >> 5, sh = bit & 31;
    if (word < F::N) r.v[word] |= l << sh;
    if (sh > 3 && word + 1 < F::N) r.v[word + 1] |= l >> (32 - sh);
  }
  fe_reduce_once(r);
  Fe<F> k;
#pragma unroll
  for (int i = 0; i < F::N; i++) k.v[i] = F::FROM29(i);
  return fe_mul(r, k);
}

template <class F>
__device__ __forceinline__ Fe<F> fe_from_le_bytes(const uint8_t* __restrict__ b, int nbytes) {
  Fe<F> r = fe_zero<F>();
  for (int i = 0; i < nbytes; i++) r.v[i >> 2] |= (uint32_t)b[i] << (8 * (i & 3));
  // canonical input expected; out-of-range values are brought into [0,p) when only one p too large
  fe_reduce_once(r);
  return fe_to_mont(r);
}

template <class F>
__device__ __forceinline__ Fe<F> fe_beta() {
  Fe<F> k;
#pragma unroll
  for (int i = 0; i < F::N; i++) k.v[i] = F::BETA(i);
  return k;
}

// One thread per input point: writes records 2i (G) and 2i+1 (endo(G)).
template <class F>
__global__ void k_ingest_points(const uint8_t* __restrict__ in, size_t n, int layout, uint4* __restrict__ bases) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Aff<F> P;
  if (layout == 0) {  // LIMB29_MONT
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + i * (2 * F::N29 + 1);
    P.x = fe_from_limb29<F>(w);
    P.y = fe_from_limb29<F>(w + F::N29);
    bool nonzero = (w[2 * F::N29] & 0xFFu) != 0;
    if (!nonzero) P = aff_inf<F>();
  } else {
    int nb = (F::BITS + 7) / 8;
    const uint8_t* b = in + i * (size_t)(2 * nb);
    P.x = fe_from_le_bytes<F>(b, nb);
    P.y = fe_from_le_bytes<F>(b + nb, nb);
  }
  Aff<F> E = P;
  if (!aff_is_inf(P)) E.x = fe_mul(P.x, fe_beta<F>());  // src/wasm/curve.ts:90-103
  uint4* o = bases + i * (size_t)(4 * F::N / 4);
  st_aos<F>(o, P.x);
  st_aos<F>(o + F::N / 4, P.y);
  st_aos<F>(o + 2 * F::N / 4, E.x);
  st_aos<F>(o + 3 * F::N / 4, E.y);
}

// scalars -> 8 x u32 plain integers
__device__ __forceinline__ void load_scalar(const uint8_t* __restrict__ in, size_t i, int layout, uint32_t* s) {
  if (layout == 0) {  // 9 x 29-bit limbs
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + i * 9;
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = 0;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      uint32_t l = w[j] & 0x1FFFFFFFu;
      int bit = 29 * j;
      int word = bit >> 5, sh = bit & 31;
      if (word < 8) s[word] |= l << sh;
      if (sh > 3 && word + 1 < 8) s[word + 1] |= l >> (32 - sh);
    }
  } else {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + i * 8;  // 32-byte records are 4-aligned
#pragma unroll
    for (int j = 0; j < 8; j++) s[j] = w[j];
  }
}

// One thread per scalar: GLV split into two half-scalars, magnitude in bits 0..126, sign in bit 127.
template <class G>
__global__ void k_glv(const uint8_t* __restrict__ in, size_t n, int layout, uint4* __restrict__ hs) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8], s0[4], s1[4];
  load_scalar(in, i, layout, s);
  scalar_reduce<G>(s);
  uint32_t flags = glv_decompose<G>(s, s0, s1);
  s0[3] |= (flags & 1u) << 31;
  s1[3] |= (flags >> 1) << 31;
  hs[2 * i] = make_uint4(s0[0], s0[1], s0[2], s0[3]);
  hs[2 * i + 1] = make_uint4(s1[0], s1[1], s1[2], s1[3]);
}

// ------------------------------------------------------------------------------------------
// counting sort of (bucket, index) pairs
// ------------------------------------------------------------------------------------------
struct SortArgs {
  const uint4* hs;   // S half-scalars
  size_t S;
  int c, K;
  uint32_t L;
  uint32_t* cnt;     // K*L bucket counts
  uint32_t* cursor;  // K*L running positions (scatter)
  const uint32_t* po0;  // pair offsets of round 0
  uint32_t* ent;        // sorted entries, 2 * P0 slots
  uint32_t* pairkey;    // bucket of every round-0 pair
  uint32_t* digits;     // optional dump (tests): S*K
};

template <bool SCATTER>
__global__ void k_hist_scatter(SortArgs a) {
  size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= a.S) return;
  uint4 q = a.hs[h];
  uint32_t s[4] = {q.x, q.y, q.z, q.w & 0x7FFFFFFFu};
  uint32_t sign = q.w >> 31;
  uint32_t carry = 0;
  for (int k = 0; k < a.K; k++) {
    uint32_t l = signed_digit<4>(s, k, a.c, carry);
    if (a.digits && !SCATTER) a.digits[h * a.K + k] = l | ((l ? (carry ^ sign) : 0u) << 31);
    if (l == 0) continue;
    uint32_t b = (uint32_t)k * a.L + (l - 1);
    if (!SCATTER) {
      atomicAdd(&a.cnt[b], 1u);
    } else {
      uint32_t pos = atomicAdd(&a.cursor[b], 1u);
      uint32_t slot = 2u * a.po0[b] + pos;
      a.ent[slot] = (uint32_t)h | ((carry ^ sign) << 31);
      a.pairkey[slot >> 1] = b;
    }
  }
}

// Block r computes, for tree round r, the exclusive scan over buckets of
// pairs_r[b] = ceil(n_r[b] / 2), n_r[b] = ceil(cnt[b] / 2^r); block 0 also the max count.
// po[r * NB + b], totals[r], totals[MAX_ROUNDS+1] = max count, totals[MAX_ROUNDS+2] = sum of counts.
static __global__ void k_scan(const uint32_t* __restrict__ cnt, uint32_t NB, uint32_t* __restrict__ po,
                       unsigned long long* __restrict__ totals) {
  __shared__ unsigned long long sh[1024];
  __shared__ uint32_t shmax[1024];
  int r = blockIdx.x;
  int t = threadIdx.x, T = blockDim.x;
  uint32_t per = (NB + T - 1) / T;
  uint32_t b0 = min(NB, (uint32_t)t * per), b1 = min(NB, b0 + per);
  unsigned long long sum = 0, nent = 0;
  uint32_t mx = 0;
  for (uint32_t b = b0; b < b1; b++) {
    uint32_t c0 = cnt[b];
    mx = max(mx, c0);
    nent += c0;
    uint32_t n = (uint32_t)(((unsigned long long)c0 + (1ull << r) - 1) >> r);
    sum += (n + 1) >> 1;
  }
  sh[t] = sum;
  shmax[t] = mx;
  __syncthreads();
  // inclusive Hillis-Steele scan over T partial sums
  for (int off = 1; off < T; off <<= 1) {
    unsigned long long v = (t >= off) ? sh[t - off] : 0;
    uint32_t m = (t >= off) ? shmax[t - off] : 0;
    __syncthreads();
    sh[t] += v;
    shmax[t] = max(shmax[t], m);
    __syncthreads();
  }
  unsigned long long run = sh[t] - sum;
  for (uint32_t b = b0; b < b1; b++) {
    uint32_t c0 = cnt[b];
    uint32_t n = (uint32_t)(((unsigned long long)c0 + (1ull << r) - 1) >> r);
    po[(size_t)r * NB + b] = (uint32_t)run;
    run += (n + 1) >> 1;
  }
  if (t == T - 1) {
    totals[r] = sh[t];
    if (r == 0) totals[MAX_ROUNDS + 1] = shmax[t];
  }
  if (r == 0) atomicAdd(&totals[MAX_ROUNDS + 2], nent);
}

// ------------------------------------------------------------------------------------------
// one round of the pairwise tree
// ------------------------------------------------------------------------------------------
template <class F>
struct RoundArgs {
  int r;                 // round number
  size_t P;              // pairs in this round
  const uint32_t* cnt;   // bucket counts (round 0)
  const uint32_t* po_r;  // pair offsets of this round
  const uint32_t* po_n;  // pair offsets of the next round
  const uint32_t* pairkey;  // bucket of each pair
  uint32_t* pairkey_next;
  // inputs
  const uint32_t* ent;   // round 0: sorted entries
  const uint4* bases;    // round 0: base points
  ElemBuf<F> in;         // round >= 1
  ElemBuf<F> out;
  // batch inversion, level 0
  uint4* prefix;         // P elements, stride = P
  uint4* tot;            // one per thread, stride = M1
  const uint4* invtot;   // inverse of tot, same layout
  size_t M1;
};

template <class F, bool R0>
__device__ __forceinline__ void load_pair(const RoundArgs<F>& a, size_t i, Aff<F>& A, Aff<F>& B, uint32_t& b,
                                          uint32_t& j) {
  b = a.pairkey[i];
  j = (uint32_t)i - a.po_r[b];
  uint32_t n = (uint32_t)(((unsigned long long)a.cnt[b] + (1ull << a.r) - 1) >> a.r);
  bool has2 = 2 * j + 1 < n;
  if (R0) {
    uint2 e = reinterpret_cast<const uint2*>(a.ent)[i];
    A = gather_base<F>(a.bases, e.x);
    B = has2 ? gather_base<F>(a.bases, e.y) : aff_inf<F>();
  } else {
    A = a.in.load(2 * i);
    B = has2 ? a.in.load(2 * i + 1) : aff_inf<F>();
  }
}

// forward pass: exclusive prefix products of the denominators, per thread
template <class F, bool R0>
__global__ void __launch_bounds__(ACC_THREADS) k_fwd(RoundArgs<F> a) {
  size_t chunk0 = (size_t)blockIdx.x * (ACC_THREADS * ACC_B0);
  size_t gid = (size_t)blockIdx.x * ACC_THREADS + threadIdx.x;
  Fe<F> run = fe_one<F>();
#pragma unroll 1
  for (int s = 0; s < ACC_B0; s++) {
    size_t i = chunk0 + (size_t)s * ACC_THREADS + threadIdx.x;
    if (i >= a.P) break;
    Aff<F> A, B;
    uint32_t b, j;
    load_pair<F, R0>(a, i, A, B, b, j);
    Fe<F> d;
    int cs = aff_add_prepare(A, B, d);
    if (cs <= AFF_DBL) {
      st_soa<F>(a.prefix, a.P, i, run);
      run = fe_mul(run, d);
    }
  }
  st_soa<F>(a.tot, a.M1, gid, run);
}

// backward pass: individual inverses from the running inverse, then finish the additions
template <class F, bool R0>
__global__ void __launch_bounds__(ACC_THREADS) k_bwd(RoundArgs<F> a) {
  size_t chunk0 = (size_t)blockIdx.x * (ACC_THREADS * ACC_B0);
  size_t gid = (size_t)blockIdx.x * ACC_THREADS + threadIdx.x;
  Fe<F> inv = ld_soa<F>(a.invtot, a.M1, gid);
#pragma unroll 1
  for (int s = ACC_B0 - 1; s >= 0; s--) {
    size_t i = chunk0 + (size_t)s * ACC_THREADS + threadIdx.x;
    if (i >= a.P) continue;
    Aff<F> A, B;
    uint32_t b, j;
    load_pair<F, R0>(a, i, A, B, b, j);
    Fe<F> d;
    int cs = aff_add_prepare(A, B, d);
    Fe<F> id = inv;
    if (cs <= AFF_DBL) {
      Fe<F> pre = ld_soa<F>(a.prefix, a.P, i);
      id = fe_mul(inv, pre);
      inv = fe_mul(inv, d);
    }
    Aff<F> R = aff_add_finish(cs, A, B, id);
    size_t e = 2 * (size_t)a.po_n[b] + j;
    a.out.store(e, R);
    if (!(e & 1)) a.pairkey_next[e >> 1] = b;
  }
}

// upper levels of the product tree: plain arrays of field elements
template <class F>
__global__ void __launch_bounds__(UP_THREADS) k_up_fwd(const uint4* __restrict__ val, size_t M, uint4* __restrict__ pre,
                                                       uint4* __restrict__ tot, size_t Mn) {
  size_t chunk0 = (size_t)blockIdx.x * (UP_THREADS * UP_B1);
  size_t gid = (size_t)blockIdx.x * UP_THREADS + threadIdx.x;
  Fe<F> run = fe_one<F>();
#pragma unroll 1
  for (int s = 0; s < UP_B1; s++) {
    size_t i = chunk0 + (size_t)s * UP_THREADS + threadIdx.x;
    if (i >= M) break;
    Fe<F> v = ld_soa<F>(val, M, i);
    st_soa<F>(pre, M, i, run);
    run = fe_mul(run, v);
  }
  st_soa<F>(tot, Mn, gid, run);
}

// in place: pre[i] <- 1 / val[i]
template <class F>
__global__ void __launch_bounds__(UP_THREADS) k_up_bwd(const uint4* __restrict__ val, size_t M, uint4* __restrict__ pre,
                                                       const uint4* __restrict__ invtot, size_t Mn) {
  size_t chunk0 = (size_t)blockIdx.x * (UP_THREADS * UP_B1);
  size_t gid = (size_t)blockIdx.x * UP_THREADS + threadIdx.x;
  Fe<F> inv = ld_soa<F>(invtot, Mn, gid);
#pragma unroll 1
  for (int s = UP_B1 - 1; s >= 0; s--) {
    size_t i = chunk0 + (size_t)s * UP_THREADS + threadIdx.x;
    if (i >= M) continue;
    Fe<F> v = ld_soa<F>(val, M, i);
    Fe<F> p = ld_soa<F>(pre, M, i);
    st_soa<F>(pre, M, i, fe_mul(inv, p));
    inv = fe_mul(inv, v);
  }
}

template <class F>
__global__ void k_inv_top(const uint4* __restrict__ val, size_t M, uint4* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  st_soa<F>(out, M, i, fe_inv(ld_soa<F>(val, M, i)));
}

}  // namespace msm
