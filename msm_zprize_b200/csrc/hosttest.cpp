// Host-compiled view of the device math core (fp.cuh / ec.cuh / glv.cuh) for CPU unit tests.
// TEST SHIM ONLY: built as libmsm_b200_hostmath.so, loaded by tests/test_hostmath.py; it is not
// part of the product library and contains no MSM.
#include "constants.cuh"
#include "ec.cuh"
#include "glv.cuh"

using namespace msm;

template <class F>
static Fe<F> ld(const uint32_t* p) {
  Fe<F> r;
  for (int i = 0; i < F::N; i++) r.v[i] = p[i];
  return r;
}
template <class F>
static void st(uint32_t* p, const Fe<F>& a) {
  for (int i = 0; i < F::N; i++) p[i] = a.v[i];
}

template <class F>
static int fe_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Fe<F> x = ld<F>(a), y = ld<F>(b), r;
  switch (op) {
    case 0: r = fe_mul(x, y); break;
    case 1: r = fe_add(x, y); break;
    case 2: r = fe_sub(x, y); break;
    case 3: r = fe_inv(x); break;
    case 4: r = fe_to_mont(x); break;
    case 5: r = fe_from_mont(x); break;
    case 6: r = fe_sqr(x); break;
    case 7: r = fe_neg(x); break;
    case 8: r = fe_inv_fermat(x); break;
    case 9: r = fe_inv_plain(x); break;
    default: return -1;
  }
  st(out, r);
  return 0;
}

// op 0: complete add, 1: mixed add (Q affine = (X,Y) of q), 2: double, 3: to affine, 4-6: unreduced add / mixed / double
template <class F, uint32_t B3>
static int proj_op(int op, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  constexpr int N = F::N;
  Proj<F> P{ld<F>(p), ld<F>(p + N), ld<F>(p + 2 * N)};
  Proj<F> Q{ld<F>(q), ld<F>(q + N), ld<F>(q + 2 * N)};
  Proj<F> R;
  if (op == 0) R = proj_add<F, B3>(P, Q);
  else if (op == 1) R = proj_add_mixed<F, B3>(P, Aff<F>{Q.X, Q.Y});
  else if (op == 2) R = proj_dbl<F, B3>(P);
  else if (op >= 4 && op <= 6) {  // unreduced variants (F::LAZY, b3 = 3); op 7: the same after proj_canon
    if constexpr (F::LAZY && B3 == 3) {
      if (op == 4) R = proj_add_nr<F>(P, Q);
      else if (op == 5) R = proj_add_mixed_nr<F>(P, Aff<F>{Q.X, Q.Y});
      else R = proj_dbl_nr<F>(P);
      // contract: coordinates < 2p; report a violation instead of a value
      Proj<F> C = proj_canon(R);
      Proj<F> C2 = proj_canon(C);
      if (!fe_eq(C.X, C2.X) || !fe_eq(C.Y, C2.Y) || !fe_eq(C.Z, C2.Z)) return -2;
    } else return -1;
  }
  else if (op == 3) {
    Aff<F> A = proj_to_aff(P);
    R.X = A.x; R.Y = A.y; R.Z = fe_zero<F>();
    R.Z.v[0] = aff_is_inf(A) ? 0 : 1;
  } else return -1;
  st(out, R.X); st(out + N, R.Y); st(out + 2 * N, R.Z);
  return 0;
}

// affine add through prepare/finish with a direct inversion; flags: bit0 P inf, bit1 Q inf
template <class F>
static int aff_op(int flags, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  constexpr int N = F::N;
  Aff<F> P{ld<F>(p), ld<F>(p + N)}, Q{ld<F>(q), ld<F>(q + N)};
  if (flags & 1) P = aff_inf<F>();
  if (flags & 2) Q = aff_inf<F>();
  Fe<F> d;
  int cs = aff_add_prepare(P, Q, d);
  Aff<F> R = aff_add_finish(cs, P, Q, fe_inv(d));
  st(out, R.x); st(out + N, R.y);
  return aff_is_inf(R) ? 1 : 0;
}

template <class F>
static int ext_op(int op, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  constexpr int N = F::N;
  Ext<F> P{ld<F>(p), ld<F>(p + N), ld<F>(p + 2 * N), ld<F>(p + 3 * N)};
  Ext<F> Q{ld<F>(q), ld<F>(q + N), ld<F>(q + 2 * N), ld<F>(q + 3 * N)};
  Ext<F> R;
  if (op == 0) R = ext_add(P, Q);
  else if (op == 1 || op == 2) R = ext_add_niels(P, niels_from_xy(Q.X, Q.Y), op == 2);
  else return -1;
  st(out, R.X); st(out + N, R.Y); st(out + 2 * N, R.Z); st(out + 3 * N, R.T);
  return 0;
}

extern "C" {
int ht_fe_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  switch (field) {
    case 0: return fe_op<Bls377Fq>(op, a, b, out);
    case 1: return fe_op<PallasFp>(op, a, b, out);
    case 2: return fe_op<Bls377Fr>(op, a, b, out);
    case 3: return fe_op<Bls381Fq>(op, a, b, out);
  }
  return -1;
}
int ht_proj_op(int field, int op, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  switch (field) {
    case 0: return proj_op<Bls377Fq, 3>(op, p, q, out);
    case 1: return proj_op<PallasFp, 15>(op, p, q, out);
    case 3: return proj_op<Bls381Fq, 12>(op, p, q, out);
  }
  return -1;
}
int ht_aff_op(int field, int flags, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  switch (field) {
    case 0: return aff_op<Bls377Fq>(flags, p, q, out);
    case 1: return aff_op<PallasFp>(flags, p, q, out);
    case 3: return aff_op<Bls381Fq>(flags, p, q, out);
  }
  return -1;
}
int ht_ext_op(int op, const uint32_t* p, const uint32_t* q, uint32_t* out) {
  return ext_op<Bls377Fr>(op, p, q, out);
}
int ht_glv(int curve, const uint32_t* s, uint32_t* s0, uint32_t* s1) {
  uint32_t t[8];
  for (int i = 0; i < 8; i++) t[i] = s[i];
  if (curve == 0) { scalar_reduce<Bls377Glv>(t); return (int)glv_decompose<Bls377Glv>(t, s0, s1); }
  if (curve == 1) { scalar_reduce<PallasGlv>(t); return (int)glv_decompose<PallasGlv>(t, s0, s1); }
  if (curve == 3) { scalar_reduce<Bls381Glv>(t); return (int)glv_decompose<Bls381Glv>(t, s0, s1); }
  return -1;
}
}
