// Bucket reduction, window combination and output normalisation, generic over the curve form.
//
// sum_l l * B_l per window is computed as a hierarchy of weighted sums.  Items (R_t, X_t) stand
// for sum_t (t * R_t + X_t); one level groups g = 2^gb consecutive items:
//     R'_u = g * sum_t R_t,      X'_u = sum_t ((t - u g) R_t + X_t)
// (level 0: the items are the bucket sums with weights t+1 and X = 0).  This is the reference's
// per-chunk running sum with the `triangle + (lstart-1) * row` identity
// (src/msm-batched-affine.ts:530-571, src/msm-basic.ts:192-223) applied recursively, so every
// level is data parallel instead of one chunk per CPU thread.  The last level leaves one item per
// window whose X is the window sum; k_horner combines the windows (:310-321).
#pragma once
#include "kernels_common.cuh"
#include "inv_quad.cuh"

namespace msm {

// ------------------------------------------------------------------------------------------
// Quad-cooperative point operations for the latency-bound tail (Horner over the windows):
// four adjacent lanes hold identical copies of the point and each computes ONE of up to four
// independent field products of a formula layer, then all four exchange the results by shuffle.
// A doubling is 2 product layers instead of 8 sequential modmuls (one lone warp needs 0.62 us
// per modmul because the IMAD.WIDE pipe accepts one warp instruction per 4 cycles).
// ------------------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ void quad_mul4(const Fe<F>& a0, const Fe<F>& b0, const Fe<F>& a1, const Fe<F>& b1,
                                          const Fe<F>& a2, const Fe<F>& b2, const Fe<F>& a3, const Fe<F>& b3,
                                          Fe<F>& p0, Fe<F>& p1, Fe<F>& p2, Fe<F>& p3) {
  const int lane = threadIdx.x & 31, q = lane & 3, base = lane & ~3;
  // operand of this lane by bitwise selection (one 3-input logic op per step): ternaries on the lane number
  // were compiled into divergent branches, ~24 reconvergence regions per doubling
  const uint32_t m1 = 0u - (uint32_t)(q & 1), m2 = 0u - (uint32_t)((q >> 1) & 1);
  Fe<F> x, y;
#pragma unroll
  for (int i = 0; i < F::N; i++) {
    const uint32_t x01 = (a0.v[i] & ~m1) | (a1.v[i] & m1), x23 = (a2.v[i] & ~m1) | (a3.v[i] & m1);
    const uint32_t y01 = (b0.v[i] & ~m1) | (b1.v[i] & m1), y23 = (b2.v[i] & ~m1) | (b3.v[i] & m1);
    x.v[i] = (x01 & ~m2) | (x23 & m2);
    y.v[i] = (y01 & ~m2) | (y23 & m2);
  }
  Fe<F> p = fe_mul_call(x, y);
#pragma unroll
  for (int i = 0; i < F::N; i++) {
    p0.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 0);
    p1.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 1);
    p2.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 2);
    p3.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 3);
  }
}

// The product of this lane's (already selected) operands, then all four products of the quad to every lane
template <class F>
__device__ __forceinline__ void quad_mul_bcast(const Fe<F>& x, const Fe<F>& y, Fe<F>& p0, Fe<F>& p1, Fe<F>& p2, Fe<F>& p3) {
  const int base = (threadIdx.x & 31) & ~3;
  Fe<F> p = fe_mul_call(x, y);
#pragma unroll
  for (int i = 0; i < F::N; i++) {
    p0.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 0);
    p1.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 1);
    p2.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 2);
    p3.v[i] = __shfl_sync(0xffffffffu, p.v[i], base + 3);
  }
}
template <class F>
__device__ __forceinline__ Fe<F> fe_bitsel(uint32_t m, const Fe<F>& a, const Fe<F>& b) {  // m all ones: a, zero: b
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = (a.v[i] & m) | (b.v[i] & ~m);
  return r;
}

// ---- curve-form traits ------------------------------------------------------------------
template <class F_, uint32_t B3_>
struct WeierCurve {
  using F = F_;
  using Acc = Proj<F>;
  static constexpr int ACC_FE = 3;   // field elements per accumulator
  static constexpr int BASE_FE = 2;  // field elements per cached base point (x | y)
  static constexpr int BASE_STRIDE = 2;  // records per input point in ctx->bases (G, endo G)
  // generic bucket method (kernels_basic.cuh): fetching the next point ahead costs 24 more registers here and loses
  // (msmProjective 2^22: 43.0 -> 46.1 ms); the twisted-Edwards kernels (8 limbs) gain from it
  static constexpr bool PREFETCH_BASE = false;
  // QUAD_LAZY (BLS12-377): accumulators held in registers have coordinates < 2p and all additions below are
  // the unreduced variants (ec.cuh proj_*_nr); memory always holds canonical values (st canonicalises).
  __device__ static Acc zero() { return proj_zero<F>(); }
  __device__ static Acc add(const Acc& a, const Acc& b) {
    if constexpr (F::LAZY && B3_ == 3) return proj_add_nr<F>(a, b);
    else return proj_add<F, B3_>(a, b);
  }
  __device__ static Acc dbl(const Acc& a) {
    if constexpr (F::LAZY && B3_ == 3) return proj_dbl_nr<F>(a);
    else return proj_dbl<F, B3_>(a);
  }
  __device__ static void st(uint4* p, const Acc& P0) {
    const Acc P = canon(P0);
    st_aos<F>(p, P.X);
    st_aos<F>(p + F::N / 4, P.Y);
    st_aos<F>(p + 2 * F::N / 4, P.Z);
  }
  __device__ static Acc ld(const uint4* p) {
    Acc P;
    P.X = ld_aos<F>(p);
    P.Y = ld_aos<F>(p + F::N / 4);
    P.Z = ld_aos<F>(p + 2 * F::N / 4);
    return P;
  }
  __device__ static Acc shfl_down(const Acc& a, int d) {
    Acc r;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
      r.X.v[i] = __shfl_down_sync(0xffffffffu, a.X.v[i], d);
      r.Y.v[i] = __shfl_down_sync(0xffffffffu, a.Y.v[i], d);
      r.Z.v[i] = __shfl_down_sync(0xffffffffu, a.Z.v[i], d);
    }
    return r;
  }
  // acc +/- base point `idx` (src/curve-projective.ts addMixed / subMixed)
  __device__ static Acc add_base(const Acc& a, const uint4* __restrict__ bases, uint32_t idx, bool neg) {
    return add_cached(a, ld_base(bases, idx), neg);
  }
  // the base point on its own, so that a loop can fetch the next one while it adds the current one
  typedef Aff<F> Base;
  __device__ static Base ld_base(const uint4* __restrict__ bases, uint32_t idx) {
    const uint4* p = bases + (size_t)idx * BASE_STRIDE * (2 * F::N / 4);
    Aff<F> Q;
    Q.x = ld_aos<F>(p);
    Q.y = ld_aos<F>(p + F::N / 4);
    return Q;
  }
  __device__ static Acc add_cached(const Acc& a, Base Q, bool neg) {
    if (aff_is_inf(Q)) return a;
    if (neg) Q.y = fe_neg(Q.y);
    if constexpr (F::LAZY && B3_ == 3) return proj_add_mixed_nr<F>(a, Q);
    else return proj_add_mixed<F, B3_>(a, Q);
  }
  // quad-cooperative doubling / addition (same formulas as proj_dbl / proj_add, RCB16 alg. 9 / 7)
  // Unreduced variants for F::LAZY fields with b3 = 3 (BLS12-377): coordinates are kept < 2p between
  // operations and every product below multiplies operands of at most 10p x 4p, 9p x 3p (< 152 p^2 / p).
  __device__ static Acc dbl_quad_lazy(const Acc& P) {
    Fe<F> t0, t1, t2, xy;
    // lane q multiplies (Y,Y) (Y,Z) (Z,Z) (X,Y): three selections per word instead of the generic six
    const int q = threadIdx.x & 3;
    const uint32_t m3 = 0u - (uint32_t)(q == 3), m2 = 0u - (uint32_t)(q == 2), m12 = 0u - (uint32_t)(q == 1 || q == 2);
    quad_mul_bcast<F>(fe_bitsel(m3, P.X, fe_bitsel(m2, P.Z, P.Y)), fe_bitsel(m12, P.Z, P.Y), t0, t1, t2, xy);  // < p
    Fe<F> z3 = fe_shl_nr<F, 3>(t0);                      // 8 t0          < 8p
    Fe<F> t2b = fe_add_nr(fe_shl_nr<F, 1>(t2), t2);      // b3 t2 = 3 t2  < 3p
    Fe<F> y3 = fe_add_nr(t0, t2b);                       //               < 4p
    Fe<F> t23 = fe_add_nr(fe_shl_nr<F, 1>(t2b), t2b);    // 3 * (3 t2)    < 9p
    Fe<F> t0b = fe_sub_nr<F, 9>(t0, t23);                // t0 - t23 + 9p < 10p
    Fe<F> x3, zz, yy, xx;
    // (z3,t2b) (z3,t1) (t0b,y3) (t0b,xy)
    const uint32_t mhi = 0u - (uint32_t)(q >= 2), m1 = 0u - (uint32_t)(q & 1);
    quad_mul_bcast<F>(fe_bitsel(mhi, t0b, z3), fe_bitsel(mhi, fe_bitsel(m1, xy, y3), fe_bitsel(m1, t1, t2b)), x3, zz, yy, xx);
    Acc R;
    R.X = fe_shl_nr<F, 1>(xx);   // < 2p
    R.Y = fe_add_nr(x3, yy);     // < 2p
    R.Z = zz;                    // < p
    return R;
  }
  __device__ static Acc add_quad_lazy(const Acc& P, const Acc& Q) {
    Fe<F> t0, t1, t2, t3, t4, y3, d0, d1;
    quad_mul4<F>(P.X, Q.X, P.Y, Q.Y, P.Z, Q.Z, fe_add_nr(P.X, P.Y), fe_add_nr(Q.X, Q.Y), t0, t1, t2, t3);  // 4p x 4p
    quad_mul4<F>(fe_add_nr(P.Y, P.Z), fe_add_nr(Q.Y, Q.Z), fe_add_nr(P.X, P.Z), fe_add_nr(Q.X, Q.Z), P.X, P.X, P.X,
                 P.X, t4, y3, d0, d1);
    t3 = fe_sub_nr<F, 2>(t3, fe_add_nr(t0, t1));          // < 3p
    t4 = fe_sub_nr<F, 2>(t4, fe_add_nr(t1, t2));          // < 3p
    y3 = fe_sub_nr<F, 2>(y3, fe_add_nr(t0, t2));          // < 3p
    t0 = fe_add_nr(fe_shl_nr<F, 1>(t0), t0);              // 3 t0 < 3p
    t2 = fe_add_nr(fe_shl_nr<F, 1>(t2), t2);              // b3 t2 < 3p
    Fe<F> z3 = fe_add_nr(t1, t2);                         // < 4p
    t1 = fe_sub_nr<F, 3>(t1, t2);                         // < 4p
    y3 = fe_add_nr(fe_shl_nr<F, 1>(y3), y3);              // b3 y3 < 9p
    Fe<F> a, b, c, d, e, f;
    quad_mul4<F>(t1, t3, y3, t4, t1, z3, y3, t0, a, b, c, d);   // 4x3, 9x3, 4x4, 9x3
    quad_mul4<F>(z3, t4, t0, t3, t0, t0, t0, t0, e, f, d0, d1);
    Acc R;
    R.X = fe_sub_nr<F, 1>(a, b);  // < 2p
    R.Y = fe_add_nr(c, d);        // < 2p
    R.Z = fe_add_nr(e, f);        // < 2p
    return R;
  }
  // canonical representative of a point whose coordinates are < 2p (before it is stored or compared)
  __device__ static Acc canon(const Acc& P) {
    Acc R = P;
    if (QUAD_LAZY) {
      fe_reduce_once(R.X);
      fe_reduce_once(R.Y);
      fe_reduce_once(R.Z);
    }
    return R;
  }
  static constexpr bool QUAD_LAZY = F::LAZY && B3_ == 3;
  // quad ops on values < 2p with results < 2p when QUAD_LAZY (canon() before storing), canonical otherwise
  __device__ static Acc dblq(const Acc& P) { return QUAD_LAZY ? dbl_quad_lazy(P) : dbl_quad(P); }
  __device__ static Acc addq(const Acc& P, const Acc& Q) { return QUAD_LAZY ? add_quad_lazy(P, Q) : add_quad(P, Q); }
  __device__ static Acc dbl_quad(const Acc& P) {
    Fe<F> t0, t1, t2, xy;
    quad_mul4<F>(P.Y, P.Y, P.Y, P.Z, P.Z, P.Z, P.X, P.Y, t0, t1, t2, xy);
    Fe<F> z3 = fe_dbl(fe_dbl(fe_dbl(t0)));
    t2 = fe_mul_small(t2, B3_);
    Fe<F> y3 = fe_add(t0, t2);
    Fe<F> t23 = fe_add(fe_dbl(t2), t2);
    Fe<F> t0b = fe_sub(t0, t23);
    Fe<F> x3, zz, yy, xx;
    quad_mul4<F>(t2, z3, t1, z3, t0b, y3, t0b, xy, x3, zz, yy, xx);
    Acc R;
    R.X = fe_dbl(xx);
    R.Y = fe_add(x3, yy);
    R.Z = zz;
    return R;
  }
  __device__ static Acc add_quad(const Acc& P, const Acc& Q) {
    Fe<F> t0, t1, t2, t3, t4, y3, d0, d1;
    quad_mul4<F>(P.X, Q.X, P.Y, Q.Y, P.Z, Q.Z, fe_add(P.X, P.Y), fe_add(Q.X, Q.Y), t0, t1, t2, t3);
    quad_mul4<F>(fe_add(P.Y, P.Z), fe_add(Q.Y, Q.Z), fe_add(P.X, P.Z), fe_add(Q.X, Q.Z), P.X, P.X, P.X, P.X, t4, y3,
                 d0, d1);
    t3 = fe_sub(t3, fe_add(t0, t1));
    t4 = fe_sub(t4, fe_add(t1, t2));
    y3 = fe_sub(y3, fe_add(t0, t2));
    t0 = fe_add(fe_dbl(t0), t0);
    t2 = fe_mul_small(t2, B3_);
    Fe<F> z3 = fe_add(t1, t2);
    t1 = fe_sub(t1, t2);
    y3 = fe_mul_small(y3, B3_);
    Fe<F> a, b, c, d, e, f;
    quad_mul4<F>(t3, t1, t4, y3, t1, z3, y3, t0, a, b, c, d);
    quad_mul4<F>(z3, t4, t0, t3, t0, t0, t0, t0, e, f, d0, d1);
    Acc R;
    R.X = fe_sub(a, b);
    R.Y = fe_add(c, d);
    R.Z = fe_add(e, f);
    return R;
  }
  // canonical affine output words: x | y | is_zero
  // `zi` = 1 / a.Z computed by the caller (quad-cooperative inversion, all lanes); a canonical
  __device__ static void normalise(const Acc& a, const Fe<F>& zi, uint32_t* out) {
    bool inf = fe_is_zero(a.Z);
    Fe<F> x = inf ? fe_zero<F>() : fe_from_mont(fe_mul(a.X, zi));
    Fe<F> y = inf ? fe_zero<F>() : fe_from_mont(fe_mul(a.Y, zi));
    for (int i = 0; i < F::N; i++) {
      out[i] = x.v[i];
      out[F::N + i] = y.v[i];
    }
    out[2 * F::N] = inf ? 1u : 0u;
  }
};

template <class F_>
struct TeCurve {
  using F = F_;
  using Acc = Ext<F>;
  static constexpr int ACC_FE = 4;
  static constexpr int BASE_FE = 3;  // (y+x | y-x | 2d*x*y)
  static constexpr int BASE_STRIDE = 1;
  static constexpr bool PREFETCH_BASE = true;
  __device__ static Acc zero() { return ext_zero<F>(); }
  __device__ static Acc add(const Acc& a, const Acc& b) { return ext_add<F>(a, b); }
  __device__ static Acc dbl(const Acc& a) { return ext_dbl<F>(a); }
  __device__ static void st(uint4* p, const Acc& P) {
    st_aos<F>(p, P.X);
    st_aos<F>(p + F::N / 4, P.Y);
    st_aos<F>(p + 2 * F::N / 4, P.Z);
    st_aos<F>(p + 3 * F::N / 4, P.T);
  }
  __device__ static Acc ld(const uint4* p) {
    Acc P;
    P.X = ld_aos<F>(p);
    P.Y = ld_aos<F>(p + F::N / 4);
    P.Z = ld_aos<F>(p + 2 * F::N / 4);
    P.T = ld_aos<F>(p + 3 * F::N / 4);
    return P;
  }
  __device__ static Acc shfl_down(const Acc& a, int d) {
    Acc r;
#pragma unroll
    for (int i = 0; i < F::N; i++) {
      r.X.v[i] = __shfl_down_sync(0xffffffffu, a.X.v[i], d);
      r.Y.v[i] = __shfl_down_sync(0xffffffffu, a.Y.v[i], d);
      r.Z.v[i] = __shfl_down_sync(0xffffffffu, a.Z.v[i], d);
      r.T.v[i] = __shfl_down_sync(0xffffffffu, a.T.v[i], d);
    }
    return r;
  }
  __device__ static Acc add_base(const Acc& a, const uint4* __restrict__ bases, uint32_t idx, bool neg) {
    return ext_add_niels<F>(a, ld_base(bases, idx), neg);
  }
  // the cached base point on its own, so that a loop can fetch the next one while it adds the current one
  typedef Niels<F> Base;
  __device__ static Base ld_base(const uint4* __restrict__ bases, uint32_t idx) {
    const uint4* p = bases + (size_t)idx * (3 * F::N / 4);
    Niels<F> Q;
    Q.yp = ld_aos<F>(p);
    Q.ym = ld_aos<F>(p + F::N / 4);
    Q.kt = ld_aos<F>(p + 2 * F::N / 4);
    return Q;
  }
  __device__ static Acc add_cached(const Acc& a, const Base& Q, bool neg) { return ext_add_niels<F>(a, Q, neg); }
  // quad-cooperative doubling (dedicated a = -1 formula dbl-2008-hwcd: 4M + 4S in 2 layers; the
  // reference doubles with the unified addition, src/curve-twisted-edwards.ts:215-227 -- same point)
  __device__ static Acc dblq(const Acc& P) { return dbl_quad(P); }
  __device__ static Acc addq(const Acc& P, const Acc& Q) { return add_quad(P, Q); }
  __device__ static Acc canon(const Acc& P) { return P; }
  __device__ static Acc dbl_quad(const Acc& P) {
    Fe<F> A, B, ZZ, S;
    quad_mul4<F>(P.X, P.X, P.Y, P.Y, P.Z, P.Z, fe_add(P.X, P.Y), fe_add(P.X, P.Y), A, B, ZZ, S);
    Fe<F> Cc = fe_dbl(ZZ);
    Fe<F> D = fe_neg(A);
    Fe<F> E = fe_sub(fe_sub(S, A), B);
    Fe<F> G = fe_add(D, B);
    Fe<F> Fv = fe_sub(G, Cc);
    Fe<F> H = fe_sub(D, B);
    Acc R;
    quad_mul4<F>(E, Fv, G, H, E, H, Fv, G, R.X, R.Y, R.T, R.Z);
    return R;
  }
  __device__ static Acc add_quad(const Acc& P, const Acc& Q) {
    Fe<F> A, B, TT, ZZ, Cc, d0, d1, d2;
    quad_mul4<F>(fe_sub(P.Y, P.X), fe_sub(Q.Y, Q.X), fe_add(P.Y, P.X), fe_add(Q.Y, Q.X), P.T, Q.T, P.Z, Q.Z, A, B, TT, ZZ);
    quad_mul4<F>(TT, fe_k2d<F>(), TT, TT, TT, TT, TT, TT, Cc, d0, d1, d2);
    Fe<F> D = fe_dbl(ZZ);
    Fe<F> E = fe_sub(B, A), Fv = fe_sub(D, Cc), G = fe_add(D, Cc), H = fe_add(B, A);
    Acc R;
    quad_mul4<F>(E, Fv, G, H, E, H, Fv, G, R.X, R.Y, R.T, R.Z);
    return R;
  }
  // (X/Z, Y/Z): src/bigint/twisted-edwards.ts:39-45; is_zero flags the neutral point (0, 1)
  __device__ static void normalise(const Acc& a, const Fe<F>& zi, uint32_t* out) {
    Fe<F> x = fe_from_mont(fe_mul(a.X, zi));
    Fe<F> y = fe_from_mont(fe_mul(a.Y, zi));
    bool zero = fe_is_zero(x);
    for (int i = 0; i < F::N; i++) {
      out[i] = x.v[i];
      out[F::N + i] = y.v[i];
      if (y.v[i] != (i == 0 ? 1u : 0u)) zero = false;
    }
    out[2 * F::N] = zero ? 1u : 0u;
  }
};

// ---- level-0 loaders ----------------------------------------------------------------------
// affine bucket sums left in the `fin` array by the batched-affine tree
template <class F, uint32_t B3>
struct AffineBucketLoader {
  const uint4* fin;  // x chunks, then y chunks, `cap` elements each
  size_t cap;
  const uint32_t* cnt;
  __device__ __forceinline__ void add_bucket(Proj<F>& run, uint32_t b) const {
    if (cnt[b] == 0) return;
    Aff<F> A;
    A.x = ld_soa<F>(fin, cap, b);
    A.y = ld_soa<F>(fin + (size_t)(F::N / 4) * cap, cap, b);
    if (aff_is_inf(A)) return;
    if constexpr (F::LAZY && B3 == 3) run = proj_add_mixed_nr<F>(run, A);
    else run = proj_add_mixed<F, B3>(run, A);
  }
  __device__ __forceinline__ Proj<F> load(uint32_t b) const {  // the bucket sum itself (neutral element if empty)
    if (cnt[b] == 0) return proj_zero<F>();
    Aff<F> A;
    A.x = ld_soa<F>(fin, cap, b);
    A.y = ld_soa<F>(fin + (size_t)(F::N / 4) * cap, cap, b);
    return proj_from_aff(A);
  }
};

// accumulators written by k_bucket_acc (basic bucket method)
template <class C>
struct AccBucketLoader {
  const uint4* buckets;
  __device__ __forceinline__ void add_bucket(typename C::Acc& run, uint32_t b) const {
    run = C::add(run, C::ld(buckets + (size_t)b * (C::ACC_FE * C::F::N / 4)));
  }
  __device__ __forceinline__ typename C::Acc load(uint32_t b) const {
    return C::ld(buckets + (size_t)b * (C::ACC_FE * C::F::N / 4));
  }
};

template <class C>
__host__ __device__ constexpr int item_u4() {
  return 2 * C::ACC_FE * C::F::N / 4;
}

// Level 0: one thread per group of g = 2^gb consecutive buckets (weights j + 1):
//   run = sum B_j,  tri = sum (j + 1) B_j   ->  item (R = g * run, X = tri).
// (A lane-quad version of this level was measured slower: the level is throughput-bound, 2 complete
// additions per bucket, and quads only help latency.)
template <class C, class Loader>
__global__ void __launch_bounds__(64) k_reduce0(Loader ld, uint32_t NB, int gb, uint4* __restrict__ out) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = 1u << gb;
  if ((size_t)u * g >= NB) return;
  typename C::Acc run = C::zero(), tri = C::zero();
#pragma unroll 1
  for (int jj = (int)g - 1; jj >= 0; jj--) {
    ld.add_bucket(run, u * g + jj);
    tri = C::add(tri, run);
  }
#pragma unroll 1
  for (int d = 0; d < gb; d++) run = C::dbl(run);
  uint4* o = out + (size_t)u * item_u4<C>();
  C::st(o, run);
  C::st(o + item_u4<C>() / 2, tri);
}

// Level 0 with one group per lane QUAD (the four lanes share the field products of every addition): for few
// buckets (shared-bucket mode, 2^15 of them) the level is bound by the latency of its 2 g dependent
// additions, not by throughput, and a quad-cooperative addition takes about half the time of a lone lane's.
template <class C, class Loader>
__global__ void __launch_bounds__(64) k_reduce0_quad(Loader ld, uint32_t NB, int gb, uint4* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t u = t >> 2;
  const uint32_t g = 1u << gb;
  const bool live = (size_t)u * g < NB;  // NB is a multiple of g; dead quads only exist in the last warp
  typename C::Acc run = C::zero(), tri = C::zero();
#pragma unroll 1
  for (int jj = (int)g - 1; jj >= 0; jj--) {
    typename C::Acc B = live ? ld.load(u * g + jj) : C::zero();
    run = C::addq(run, B);
    tri = C::addq(tri, run);
  }
#pragma unroll 1
  for (int d = 0; d < gb; d++) run = C::dblq(run);
  if (live && (t & 3) == 0) {
    uint4* o = out + (size_t)u * item_u4<C>();
    C::st(o, run);
    C::st(o + item_u4<C>() / 2, tri);
  }
}

// Levels >= 1 with one item per lane: groups of g = 2^gb consecutive lanes (g <= 32).
//   S_j = sum_{t >= j} R_t   (suffix scan over the group, gb shuffle steps)
//   X'  = sum_j X_j + sum_{j >= 1} S_j   (one add + gb-step tree reduction),  R' = g * S_0
// Serial depth gb + 1 + gb additions for gb bits of the window instead of 3 * 2^gb.
template <class C>
__global__ void __launch_bounds__(64) k_reduce_warp(const uint4* __restrict__ in, uint32_t n_items, int gb,
                                                    uint4* __restrict__ out) {
  typedef typename C::Acc Acc;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = 1 << gb;
  const int j = (int)(i & (uint32_t)(g - 1));
  const bool live = i < n_items;  // n_items is a multiple of g, so whole groups are live or dead
  const uint4* p = in + (size_t)(live ? i : 0) * item_u4<C>();
  Acc S = live ? C::ld(p) : C::zero();
  Acc Y = live ? C::ld(p + item_u4<C>() / 2) : C::zero();
#pragma unroll 1
  for (int d = 1; d < g; d <<= 1) {
    Acc t = C::shfl_down(S, d);
    Acc n = C::add(S, t);
    if (j + d < g) S = n;
  }
  {
    Acc n = C::add(Y, S);
    if (j >= 1) Y = n;
  }
#pragma unroll 1
  for (int d = g >> 1; d >= 1; d >>= 1) {
    Acc t = C::shfl_down(Y, d);
    Acc n = C::add(Y, t);
    if (j < d) Y = n;
  }
#pragma unroll 1
  for (int d = 0; d < gb; d++) S = C::dbl(S);
  if (live && j == 0) {
    uint4* o = out + (size_t)(i >> gb) * item_u4<C>();
    C::st(o, S);
    C::st(o + item_u4<C>() / 2, Y);
  }
}

// Same weighted-sum level as k_reduce_warp, but every item is held by a lane QUAD that shares the
// field products of each complete addition (latency-bound levels with few items: 3 bits per level,
// (3 + 1 + 3) additions of ~4.5 us instead of (5 + 1 + 5) of ~8.7 us).
template <class C>
__global__ void __launch_bounds__(64) k_reduce_quad(const uint4* __restrict__ in, uint32_t n_items, int gb,
                                                    uint4* __restrict__ out) {
  typedef typename C::Acc Acc;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = t >> 2;  // item
  const int g = 1 << gb;      // items per group, g <= 8 (a group spans 4 g <= 32 lanes)
  const int j = (int)(i & (uint32_t)(g - 1));
  const bool live = i < n_items;
  const uint4* p = in + (size_t)(live ? i : 0) * item_u4<C>();
  Acc S = live ? C::ld(p) : C::zero();
  Acc Y = live ? C::ld(p + item_u4<C>() / 2) : C::zero();
#pragma unroll 1
  for (int d = 1; d < g; d <<= 1) {
    Acc n = C::addq(S, C::shfl_down(S, 4 * d));
    if (j + d < g) S = n;
  }
  {
    Acc n = C::addq(Y, S);
    if (j >= 1) Y = n;
  }
#pragma unroll 1
  for (int d = g >> 1; d >= 1; d >>= 1) {
    Acc n = C::addq(Y, C::shfl_down(Y, 4 * d));
    if (j < d) Y = n;
  }
#pragma unroll 1
  for (int d = 0; d < gb; d++) S = C::dblq(S);
  if (live && j == 0 && (t & 3) == 0) {
    uint4* o = out + (size_t)(i >> gb) * item_u4<C>();
    C::st(o, S);
    C::st(o + item_u4<C>() / 2, Y);
  }
}

// One warp; every lane quad runs the same Horner chain cooperatively (src/msm-batched-affine.ts:310-321).
template <class C>
__global__ void __launch_bounds__(32) k_horner(const uint4* __restrict__ items, int K, int c, uint4* __restrict__ partial) {
  typename C::Acc acc = C::ld(items + (size_t)(K - 1) * item_u4<C>() + item_u4<C>() / 2);
#pragma unroll 1
  for (int k = K - 2; k >= 0; k--) {
#pragma unroll 1
    for (int d = 0; d < c; d++) acc = C::dblq(acc);
    acc = C::addq(acc, C::ld(items + (size_t)k * item_u4<C>() + item_u4<C>() / 2));
  }
  if (threadIdx.x == 0) C::st(partial, acc);
}

template <class C>
__global__ void k_zero_partial(uint4* __restrict__ partial) {
  if (threadIdx.x == 0 && blockIdx.x == 0) C::st(partial, C::zero());
}

// Sum `count` partials and normalise (Projective.toAffine + fromMontgomery,
// src/curve-projective.ts:335-349, src/field-msm.ts:182-185).
template <class C>
__global__ void k_finalize(const uint4* __restrict__ partials, int count, uint32_t* __restrict__ out) {
  // one warp: every lane quad sums the partials with shared field products (add_quad), lane 0 normalises
  typename C::Acc acc = C::ld(partials);
#pragma unroll 1
  for (int i = 1; i < count; i++) acc = C::addq(acc, C::ld(partials + (size_t)i * (C::ACC_FE * C::F::N / 4)));
  acc = C::canon(acc);
  const Fe<typename C::F> zi = fe_inv_quad(acc.Z);  // every lane holds the same point: the warp inverts together
  if (threadIdx.x == 0 && blockIdx.x == 0) C::normalise(acc, zi, out);
}

}  // namespace msm
