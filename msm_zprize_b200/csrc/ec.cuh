// Curve arithmetic on top of fp.cuh.
//
//  * short Weierstrass a = 0, affine with batched inversion: the pieces used by the accumulation
//    kernels (replaces batchAddNew / batchAddUnsafeNew, src/curve-affine.ts:376-522, and wasm
//    addAffine, src/wasm/curve.ts:32-84)
//  * short Weierstrass a = 0, homogeneous projective, COMPLETE formulas (Renes-Costello-Batina
//    2016, algorithms 7-9 for a = 0) for the bucket reduction / Horner / final sum.  The
//    reference uses add-1998-cmo-2 + explicit zero/double branches (src/curve-projective.ts:51-160,
//    202-253); the group element computed is the same, only the representative differs, and only
//    the normalised affine output is compared.
//  * twisted Edwards a = -1 extended coordinates, add-2008-hwcd-3 with k = 2d
//    (src/curve-twisted-edwards.ts:84-165), with input points cached as (y+x, y-x, 2d*x*y).
#pragma once
#include "fp.cuh"

namespace msm {

// ------------------------------------------------------------------------------------------
// affine points; infinity is flagged by an impossible top limb of x
// ------------------------------------------------------------------------------------------
template <class F>
struct Aff {
  Fe<F> x, y;
};

constexpr uint32_t AFF_INF_MARK = 0xFFFFFFFFu;

template <class F>
MSM_HD bool aff_is_inf(const Aff<F>& P) {
  return P.x.v[F::N - 1] == AFF_INF_MARK;
}

template <class F>
MSM_HD Aff<F> aff_inf() {
  Aff<F> P;
  P.x = fe_zero<F>();
  P.y = fe_zero<F>();
  P.x.v[F::N - 1] = AFF_INF_MARK;
  return P;
}

template <class F>
MSM_HD Aff<F> aff_neg(const Aff<F>& P) {
  Aff<F> R;
  R.x = P.x;
  R.y = fe_neg(P.y);
  return R;
}

// Case analysis of one affine addition P + Q (safe semantics of src/curve-affine.ts:376-458):
enum AffCase : int {
  AFF_ADD = 0,     // generic: slope (y2-y1)/(x2-x1)
  AFF_DBL = 1,     // P == Q, y != 0: slope 3x^2/(2y)
  AFF_TAKE_P = 2,  // Q is infinity (or pair has no second element): result P
  AFF_TAKE_Q = 3,  // P is infinity: result Q
  AFF_ZERO = 4,    // P == -Q (or doubling a point with y == 0): result infinity
};

// Returns the case and the denominator that goes into the batch inversion (1 when unused).
template <class F>
MSM_HD int aff_add_prepare(const Aff<F>& P, const Aff<F>& Q, Fe<F>& denom) {
  denom = fe_one<F>();
  if (aff_is_inf(Q)) return AFF_TAKE_P;
  if (aff_is_inf(P)) return AFF_TAKE_Q;
  Fe<F> dx = fe_sub(Q.x, P.x);
  if (fe_is_zero(dx)) {
    if (fe_eq(P.y, Q.y) && !fe_is_zero(P.y)) {
      denom = fe_dbl(P.y);
      return AFF_DBL;
    }
    return AFF_ZERO;
  }
  denom = dx;
  return AFF_ADD;
}

// Finishes P + Q given inv = 1/denom.  3 mul-equivalents (1S + 2M), +1S +small adds for doubling.
template <class F>
MSM_HD Aff<F> aff_add_finish(int cs, const Aff<F>& P, const Aff<F>& Q, const Fe<F>& inv) {
  if (cs == AFF_TAKE_P) return P;
  if (cs == AFF_TAKE_Q) return Q;
  if (cs == AFF_ZERO) return aff_inf<F>();
  Fe<F> num;
  if (cs == AFF_DBL) {
    Fe<F> xx = fe_sqr(P.x);
    num = fe_add(fe_dbl(xx), xx);  // 3 x^2   (a = 0)
  } else {
    num = fe_sub(Q.y, P.y);
  }
  Fe<F> m = fe_mul(num, inv);
  Aff<F> R;
  R.x = fe_sub(fe_sub(fe_sqr(m), P.x), Q.x);  // for doubling Q.x == P.x
  R.y = fe_sub(fe_mul(m, fe_sub(P.x, R.x)), P.y);
  return R;
}

// ------------------------------------------------------------------------------------------
// homogeneous projective, complete formulas for y^2 = x^3 + b   (b3 = 3b small: 3 or 15)
// ------------------------------------------------------------------------------------------
template <class F>
struct Proj {
  Fe<F> X, Y, Z;
};

template <class F>
MSM_HD Proj<F> proj_zero() {
  Proj<F> P;
  P.X = fe_zero<F>();
  P.Y = fe_one<F>();
  P.Z = fe_zero<F>();
  return P;
}

template <class F>
MSM_HD Proj<F> proj_from_aff(const Aff<F>& A) {
  if (aff_is_inf(A)) return proj_zero<F>();
  Proj<F> P;
  P.X = A.x;
  P.Y = A.y;
  P.Z = fe_one<F>();
  return P;
}

template <class F, uint32_t B3>
MSM_HD Proj<F> proj_add(const Proj<F>& P, const Proj<F>& Q) {
  Fe<F> t0 = fe_mul_call(P.X, Q.X);
  Fe<F> t1 = fe_mul_call(P.Y, Q.Y);
  Fe<F> t2 = fe_mul_call(P.Z, Q.Z);
  Fe<F> t3 = fe_mul_call(fe_add(P.X, P.Y), fe_add(Q.X, Q.Y));
  t3 = fe_sub(t3, fe_add(t0, t1));
  Fe<F> t4 = fe_mul_call(fe_add(P.Y, P.Z), fe_add(Q.Y, Q.Z));
  t4 = fe_sub(t4, fe_add(t1, t2));
  Fe<F> y3 = fe_mul_call(fe_add(P.X, P.Z), fe_add(Q.X, Q.Z));
  y3 = fe_sub(y3, fe_add(t0, t2));
  t0 = fe_add(fe_dbl(t0), t0);
  t2 = fe_mul_small(t2, B3);
  Fe<F> z3 = fe_add(t1, t2);
  t1 = fe_sub(t1, t2);
  y3 = fe_mul_small(y3, B3);
  Proj<F> R;
  R.X = fe_sub(fe_mul_call(t3, t1), fe_mul_call(t4, y3));
  R.Y = fe_add(fe_mul_call(t1, z3), fe_mul_call(y3, t0));
  R.Z = fe_add(fe_mul_call(z3, t4), fe_mul_call(t0, t3));
  return R;
}

// P + Q with Q affine and NOT infinity (callers check the flag)
template <class F, uint32_t B3>
MSM_HD Proj<F> proj_add_mixed(const Proj<F>& P, const Aff<F>& Q) {
  Fe<F> t0 = fe_mul_call(P.X, Q.x);
  Fe<F> t1 = fe_mul_call(P.Y, Q.y);
  Fe<F> t3 = fe_mul_call(fe_add(Q.x, Q.y), fe_add(P.X, P.Y));
  t3 = fe_sub(t3, fe_add(t0, t1));
  Fe<F> t4 = fe_add(fe_mul_call(Q.y, P.Z), P.Y);
  Fe<F> y3 = fe_add(fe_mul_call(Q.x, P.Z), P.X);
  t0 = fe_add(fe_dbl(t0), t0);
  Fe<F> t2 = fe_mul_small(P.Z, B3);
  Fe<F> z3 = fe_add(t1, t2);
  t1 = fe_sub(t1, t2);
  y3 = fe_mul_small(y3, B3);
  Proj<F> R;
  R.X = fe_sub(fe_mul_call(t3, t1), fe_mul_call(t4, y3));
  R.Y = fe_add(fe_mul_call(t1, z3), fe_mul_call(y3, t0));
  R.Z = fe_add(fe_mul_call(z3, t4), fe_mul_call(t0, t3));
  return R;
}

template <class F, uint32_t B3>
MSM_HD Proj<F> proj_dbl(const Proj<F>& P) {
  Fe<F> t0 = fe_sqr_call(P.Y);
  Fe<F> z3 = fe_dbl(fe_dbl(fe_dbl(t0)));  // 8 Y^2
  Fe<F> t1 = fe_mul_call(P.Y, P.Z);
  Fe<F> t2 = fe_mul_small(fe_sqr_call(P.Z), B3);
  Fe<F> x3 = fe_mul_call(t2, z3);
  Fe<F> y3 = fe_add(t0, t2);
  z3 = fe_mul_call(t1, z3);
  t2 = fe_add(fe_dbl(t2), t2);
  t0 = fe_sub(t0, t2);
  y3 = fe_add(x3, fe_mul_call(t0, y3));
  t1 = fe_mul_call(P.X, P.Y);
  Proj<F> R;
  R.X = fe_dbl(fe_mul_call(t0, t1));
  R.Y = y3;
  R.Z = z3;
  return R;
}

// (X/Z, Y/Z), infinity if Z == 0   (src/curve-projective.ts:335-349)
// ------------------------------------------------------------------------------------------
// The same three formulas with unreduced field additions, for F::LAZY fields with b3 = 3 (BLS12-377 Fq:
// floor(2^384 / p) = 152).  Contract: coordinates in AND out are < 2p (affine operands canonical); every
// product multiplies operands a < 10p (the full-width one) and b with (a/p)(b/p) <= 70 < 152, so fe_mul returns
// canonical values; differences add the multiple of p noted beside them.  Callers canonicalise (proj_canon)
// before a value is stored, compared or inverted.
// ------------------------------------------------------------------------------------------
template <class F>
MSM_HD Fe<F> fe_x3_nr(const Fe<F>& a) {  // 3a
  return fe_add_nr(fe_shl_nr<F, 1>(a), a);
}

template <class F>
MSM_HD Proj<F> proj_add_nr(const Proj<F>& P, const Proj<F>& Q) {
  Fe<F> t0 = fe_mul_call(P.X, Q.X);
  Fe<F> t1 = fe_mul_call(P.Y, Q.Y);
  Fe<F> t2 = fe_mul_call(P.Z, Q.Z);
  Fe<F> t3 = fe_mul_call(fe_add_nr(P.X, P.Y), fe_add_nr(Q.X, Q.Y));  // 4p x 4p
  t3 = fe_sub_nr<F, 2>(t3, fe_add_nr(t0, t1));                        // < 3p
  Fe<F> t4 = fe_mul_call(fe_add_nr(P.Y, P.Z), fe_add_nr(Q.Y, Q.Z));
  t4 = fe_sub_nr<F, 2>(t4, fe_add_nr(t1, t2));                        // < 3p
  Fe<F> y3 = fe_mul_call(fe_add_nr(P.X, P.Z), fe_add_nr(Q.X, Q.Z));
  y3 = fe_sub_nr<F, 2>(y3, fe_add_nr(t0, t2));                        // < 3p
  t0 = fe_x3_nr(t0);                                                  // < 3p
  t2 = fe_x3_nr(t2);                                                  // b3 t2 < 3p
  Fe<F> z3 = fe_add_nr(t1, t2);                                       // < 4p
  t1 = fe_sub_nr<F, 3>(t1, t2);                                       // < 4p
  y3 = fe_x3_nr(y3);                                                  // b3 y3 < 9p
  Proj<F> R;
  R.X = fe_sub_nr<F, 1>(fe_mul_call(t1, t3), fe_mul_call(y3, t4));    // 4x3, 9x3
  R.Y = fe_add_nr(fe_mul_call(t1, z3), fe_mul_call(y3, t0));          // 4x4, 9x3
  R.Z = fe_add_nr(fe_mul_call(z3, t4), fe_mul_call(t0, t3));          // 4x3, 3x3
  return R;
}

// Q affine, canonical, not infinity
template <class F>
MSM_HD Proj<F> proj_add_mixed_nr(const Proj<F>& P, const Aff<F>& Q) {
  Fe<F> t0 = fe_mul_call(P.X, Q.x);
  Fe<F> t1 = fe_mul_call(P.Y, Q.y);
  Fe<F> t3 = fe_mul_call(fe_add_nr(P.X, P.Y), fe_add_nr(Q.x, Q.y));   // 4p x 2p
  t3 = fe_sub_nr<F, 2>(t3, fe_add_nr(t0, t1));                         // < 3p
  Fe<F> t4 = fe_add_nr(fe_mul_call(P.Z, Q.y), P.Y);                    // < 3p
  Fe<F> y3 = fe_add_nr(fe_mul_call(P.Z, Q.x), P.X);                    // < 3p
  t0 = fe_x3_nr(t0);                                                   // < 3p
  Fe<F> t2 = fe_x3_nr(P.Z);                                            // b3 Z < 6p
  Fe<F> z3 = fe_add_nr(t1, t2);                                        // < 7p
  t1 = fe_sub_nr<F, 9>(t1, t2);                                        // < 10p
  y3 = fe_x3_nr(y3);                                                   // < 9p
  Proj<F> R;
  R.X = fe_sub_nr<F, 1>(fe_mul_call(t1, t3), fe_mul_call(y3, t4));     // 10x3, 9x3
  R.Y = fe_add_nr(fe_mul_call(t1, z3), fe_mul_call(y3, t0));           // 10x7, 9x3
  R.Z = fe_add_nr(fe_mul_call(z3, t4), fe_mul_call(t0, t3));           // 7x3, 3x3
  return R;
}

template <class F>
MSM_HD Proj<F> proj_dbl_nr(const Proj<F>& P) {
  Fe<F> t0 = fe_sqr_call(P.Y);
  Fe<F> z3 = fe_shl_nr<F, 3>(t0);                      // 8 t0 < 8p
  Fe<F> t1 = fe_mul_call(P.Y, P.Z);
  Fe<F> t2 = fe_x3_nr(fe_sqr_call(P.Z));               // b3 Z^2 < 3p
  Fe<F> x3 = fe_mul_call(z3, t2);                      // 8x3
  Fe<F> y3 = fe_add_nr(t0, t2);                        // < 4p
  Fe<F> zz = fe_mul_call(z3, t1);                      // 8x1
  Fe<F> t0b = fe_sub_nr<F, 9>(t0, fe_x3_nr(t2));       // t0 - 3 t2 + 9p < 10p
  Proj<F> R;
  R.Y = fe_add_nr(x3, fe_mul_call(t0b, y3));           // 10x4
  R.X = fe_shl_nr<F, 1>(fe_mul_call(t0b, fe_mul_call(P.X, P.Y)));
  R.Z = zz;
  return R;
}

template <class F>
MSM_HD Proj<F> proj_canon(const Proj<F>& P) {  // coordinates < 2p -> [0, p)
  Proj<F> R = P;
  fe_reduce_once(R.X);
  fe_reduce_once(R.Y);
  fe_reduce_once(R.Z);
  return R;
}

template <class F>
MSM_HD Aff<F> proj_to_aff(const Proj<F>& P) {
  if (fe_is_zero(P.Z)) return aff_inf<F>();
  Fe<F> zi = fe_inv(P.Z);
  Aff<F> A;
  A.x = fe_mul(P.X, zi);
  A.y = fe_mul(P.Y, zi);
  return A;
}

// ------------------------------------------------------------------------------------------
// twisted Edwards a = -1, extended coordinates
// ------------------------------------------------------------------------------------------
template <class F>
struct Ext {
  Fe<F> X, Y, Z, T;
};

// cached affine input point: (y+x, y-x, 2d*x*y); its negation swaps the first two and negates kt
template <class F>
struct Niels {
  Fe<F> yp, ym, kt;
};

template <class F>
MSM_HD Ext<F> ext_zero() {
  Ext<F> P;
  P.X = fe_zero<F>();
  P.Y = fe_one<F>();
  P.Z = fe_one<F>();
  P.T = fe_zero<F>();
  return P;
}

template <class F>
MSM_HD Fe<F> fe_k2d() {
  Fe<F> k;
#pragma unroll
  for (int i = 0; i < F::N; i++) k.v[i] = F::K2D(i);
  return k;
}

template <class F>
MSM_HD Niels<F> niels_from_xy(const Fe<F>& x, const Fe<F>& y) {
  Niels<F> n;
  n.yp = fe_add(y, x);
  n.ym = fe_sub(y, x);
  n.kt = fe_mul(fe_mul(x, y), fe_k2d<F>());
  return n;
}

// P + Q (full, 9M):  src/curve-twisted-edwards.ts:84-165 with mixed = false
template <class F>
MSM_HD Ext<F> ext_add(const Ext<F>& P, const Ext<F>& Q) {
  Fe<F> A = fe_mul_call(fe_sub(P.Y, P.X), fe_sub(Q.Y, Q.X));
  Fe<F> B = fe_mul_call(fe_add(P.Y, P.X), fe_add(Q.Y, Q.X));
  Fe<F> C = fe_mul_call(fe_mul_call(P.T, Q.T), fe_k2d<F>());
  Fe<F> D = fe_dbl(fe_mul_call(P.Z, Q.Z));
  Fe<F> E = fe_sub(B, A), Fv = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
  Ext<F> R;
  R.X = fe_mul_call(E, Fv);
  R.Y = fe_mul_call(G, H);
  R.T = fe_mul_call(E, H);
  R.Z = fe_mul_call(Fv, G);
  return R;
}

// P +/- Q with Q a cached affine point (7M): addMixed / subMixed, :199-208
template <class F>
MSM_HD Ext<F> ext_add_niels(const Ext<F>& P, const Niels<F>& Q, bool negate) {
  Fe<F> qm = fe_select(negate, Q.yp, Q.ym);
  Fe<F> qp = fe_select(negate, Q.ym, Q.yp);
  Fe<F> A = fe_mul(fe_sub(P.Y, P.X), qm);
  Fe<F> B = fe_mul(fe_add(P.Y, P.X), qp);
  Fe<F> C = fe_mul(P.T, Q.kt);
  if (negate) C = fe_neg(C);
  Fe<F> D = fe_dbl(P.Z);
  Fe<F> E = fe_sub(B, A), Fv = fe_sub(D, C), G = fe_add(D, C), H = fe_add(B, A);
  Ext<F> R;
  R.X = fe_mul(E, Fv);
  R.Y = fe_mul(G, H);
  R.T = fe_mul(E, H);
  R.Z = fe_mul(Fv, G);
  return R;
}

template <class F>
MSM_HD Ext<F> ext_dbl(const Ext<F>& P) {  // double = add(P, P): :215-227
  return ext_add(P, P);
}

}  // namespace msm
