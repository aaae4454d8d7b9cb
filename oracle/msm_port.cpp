// CPU port of the reference's MSM algorithms -- TEST / BASELINE INFRASTRUCTURE ONLY.
//
// A multi-threaded C++ restatement of mitschabaude/msm-zprize's hot path, used (a) as the checker
// for GPU results at sizes the python oracle cannot reach and (b) as the `cpu_baseline` /
// `--impl reference` arm of bench.py (the reference itself is TypeScript + runtime-generated wasm
// and cannot run in this image: no node).  Only tests/, __graft_entry__.smoke() and bench.py's
// CPU-baseline legs load this library; the product (libmsm_b200.so) never does.
//
// Parity status: PINNED through tests/test_port.py -- every result is compared with
// oracle/bigint_oracle.py (itself pinned to the reference's known-answer vectors).
//
// What follows the reference (paths relative to its repo):
//   Field<N>::mul           src/wasm/multiply-montgomery.ts:58-136  (w = 29 limbs in u64, lazy carries,
//                           result in [0, 2p); carries only at column 0 and at the end)
//   add/sub/subPositive     src/wasm/field-arithmetic.ts:32-166     (lazy range [0, 2p))
//   batch_add               src/curve-affine.ts:376-458 (safe) / 463-522 (Montgomery trick)
//   Proj add/dbl            src/curve-projective.ts:51-160,202-253  (add-1998-cmo-2, dbl-1998-cmo-2)
//   Te add                  src/curve-twisted-edwards.ts:84-165     (add-2008-hwcd-3, k = 2d)
//   glv_decompose           src/wasm/glv.ts:68-169
//   msm_affine              src/msm-batched-affine.ts:74-328        (all phases, static thread split)
//   msm_basic               src/msm-basic.ts:45-176
// Differences (result-invariant): inversion is Fermat instead of the Kaliski almost-inverse
// (src/wasm/inverse.ts); bucket chunks are balanced by point count instead of splitBuckets'
// weights (src/msm-common.ts:88-188); threads are std::thread instead of a worker pool.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <functional>
#include <thread>
#include <vector>

typedef unsigned __int128 u128;
static const int W = 29;
static const uint64_t MASK = (1ull << W) - 1;

// ---------------------------------------------------------------------------------------------
// tiny fixed-size big integers (little-endian u64 limbs) for constants and the GLV split
// ---------------------------------------------------------------------------------------------
struct Big {
  uint64_t v[10];
  Big() { memset(v, 0, sizeof v); }
  static Big from_le(const uint8_t* b, int n) {
    Big r;
    for (int i = 0; i < n; i++) r.v[i >> 3] |= (uint64_t)b[i] << (8 * (i & 7));
    return r;
  }
  void to_le(uint8_t* b, int n) const {
    for (int i = 0; i < n; i++) b[i] = (uint8_t)(v[i >> 3] >> (8 * (i & 7)));
  }
  bool bit(int i) const { return (v[i >> 6] >> (i & 63)) & 1; }
  int bits() const {
    for (int i = 639; i >= 0; i--)
      if (bit(i)) return i + 1;
    return 0;
  }
};
static Big big_add(const Big& a, const Big& b) {
  Big r;
  u128 c = 0;
  for (int i = 0; i < 10; i++) {
    c += (u128)a.v[i] + b.v[i];
    r.v[i] = (uint64_t)c;
    c >>= 64;
  }
  return r;
}
static Big big_sub(const Big& a, const Big& b) {  // mod 2^640
  Big r;
  uint64_t br = 0;
  for (int i = 0; i < 10; i++) {
    u128 t = (u128)a.v[i] - b.v[i] - br;
    r.v[i] = (uint64_t)t;
    br = (uint64_t)(t >> 64) & 1;
  }
  return r;
}
static Big big_mul(const Big& a, const Big& b) {  // low 640 bits
  Big r;
  for (int i = 0; i < 10; i++) {
    u128 c = 0;
    for (int j = 0; i + j < 10; j++) {
      c += (u128)a.v[i] * b.v[j] + r.v[i + j];
      r.v[i + j] = (uint64_t)c;
      c >>= 64;
    }
  }
  return r;
}
static Big big_shr(const Big& a, int s) {
  Big r;
  int w = s >> 6, b = s & 63;
  for (int i = 0; i + w < 10; i++) {
    r.v[i] = a.v[i + w] >> b;
    if (b && i + w + 1 < 10) r.v[i] |= a.v[i + w + 1] << (64 - b);
  }
  return r;
}
static int big_cmp(const Big& a, const Big& b) {
  for (int i = 9; i >= 0; i--) {
    if (a.v[i] > b.v[i]) return 1;
    if (a.v[i] < b.v[i]) return -1;
  }
  return 0;
}
static bool big_neg_flag(const Big& a) { return a.v[9] >> 63; }
static Big big_negate(const Big& a) { return big_sub(Big(), a); }

// ---------------------------------------------------------------------------------------------
// prime field, N limbs of 29 bits in u32 words, Montgomery radix 2^(29N), lazy range [0, 2p)
// ---------------------------------------------------------------------------------------------
template <int N>
struct Field {
  uint32_t p[N], p2[N], one[N], r2[N];
  uint64_t mu;  // -p^-1 mod 2^29
  Big pbig;

  static void from_big(uint32_t* x, const Big& b) {
    for (int i = 0; i < N; i++) {
      int bit = W * i;
      uint64_t lo = b.v[bit >> 6] >> (bit & 63);
      if ((bit & 63) + W > 64 && (bit >> 6) + 1 < 10) lo |= b.v[(bit >> 6) + 1] << (64 - (bit & 63));
      x[i] = (uint32_t)(lo & MASK);
    }
  }
  static Big to_big(const uint32_t* x) {
    Big b;
    for (int i = 0; i < N; i++) {
      int bit = W * i;
      b.v[bit >> 6] |= (uint64_t)x[i] << (bit & 63);
      if ((bit & 63) + W > 64) b.v[(bit >> 6) + 1] |= (uint64_t)x[i] >> (64 - (bit & 63));
    }
    return b;
  }
  void init(const Big& pb) {
    pbig = pb;
    from_big(p, pb);
    from_big(p2, big_add(pb, pb));
    uint64_t inv = 1;  // Newton: p^-1 mod 2^29
    for (int i = 0; i < 6; i++) inv = (inv * (2 - p[0] * inv)) & MASK;
    mu = (MASK + 1 - inv) & MASK;
    // one = 2^(29N) mod p, r2 = 2^(2*29N) mod p by repeated doubling
    Big x;
    x.v[0] = 1;
    for (int i = 0; i < 2 * W * N; i++) {
      x = big_add(x, x);
      if (big_cmp(x, pb) >= 0) x = big_sub(x, pb);
      if (i == W * N - 1) from_big(one, x);
    }
    from_big(r2, x);
  }
  static bool geq(const uint32_t* x, const uint32_t* y) {
    for (int i = N - 1; i >= 0; i--) {
      if (x[i] > y[i]) return true;
      if (x[i] < y[i]) return false;
    }
    return true;
  }
  static void sub_raw(uint32_t* z, const uint32_t* x, const uint32_t* y) {  // x >= y
    int64_t c = 0;
    for (int i = 0; i < N; i++) {
      c += (int64_t)x[i] - y[i];
      z[i] = (uint32_t)(c & MASK);
      c >>= W;
    }
  }
  static void add_raw(uint32_t* z, const uint32_t* x, const uint32_t* y) {
    uint64_t c = 0;
    for (int i = 0; i < N; i++) {
      c += (uint64_t)x[i] + y[i];
      z[i] = (uint32_t)(c & MASK);
      c >>= W;
    }
  }
  // src/wasm/multiply-montgomery.ts:58-136
  void mul(uint32_t* z, const uint32_t* x, const uint32_t* y) const {
    uint64_t S[N + 1];
    for (int j = 0; j <= N; j++) S[j] = 0;
    for (int i = 0; i < N; i++) {
      uint64_t xi = x[i];
      uint64_t t = S[0] + xi * y[0];
      uint64_t q = ((t & MASK) * mu) & MASK;
      uint64_t carry = (t + q * p[0]) >> W;
      for (int j = 1; j < N; j++) S[j - 1] = S[j] + xi * y[j] + q * p[j];
      S[0] += carry;
      S[N - 1] = 0;
    }
    uint64_t c = 0;
    for (int j = 0; j < N; j++) {
      c += S[j];
      z[j] = (uint32_t)(c & MASK);
      c >>= W;
    }
  }
  void sqr(uint32_t* z, const uint32_t* x) const { mul(z, x, x); }
  // src/wasm/field-arithmetic.ts:32-63: x + y, minus 2p if that does not underflow
  void add(uint32_t* z, const uint32_t* x, const uint32_t* y) const {
    add_raw(z, x, y);
    if (geq(z, p2)) sub_raw(z, z, p2);
  }
  // :65-100: x - y, plus 2p if negative
  void sub(uint32_t* z, const uint32_t* x, const uint32_t* y) const {
    if (geq(x, y)) {
      sub_raw(z, x, y);
    } else {
      uint32_t t[N];
      add_raw(t, x, p2);
      sub_raw(z, t, y);
    }
  }
  // :150-166 reduce to [0, p)
  void reduce(uint32_t* x) const {
    if (geq(x, p)) sub_raw(x, x, p);
    if (geq(x, p)) sub_raw(x, x, p);
  }
  bool is_zero(const uint32_t* x) const {
    uint32_t t[N];
    memcpy(t, x, sizeof t);
    reduce(t);
    for (int i = 0; i < N; i++)
      if (t[i]) return false;
    return true;
  }
  bool is_equal(const uint32_t* x, const uint32_t* y) const {
    uint32_t t[N];
    sub(t, x, y);
    return is_zero(t);
  }
  void inverse(uint32_t* z, const uint32_t* x) const {  // x^(p-2), Montgomery domain
    Big e = pbig;
    e.v[0] -= 2;  // p is odd and > 2
    uint32_t r[N], b[N];
    memcpy(r, one, sizeof r);
    memcpy(b, x, sizeof b);
    int nb = e.bits();
    for (int i = nb - 1; i >= 0; i--) {
      sqr(r, r);
      if (e.bit(i)) mul(r, r, b);
    }
    memcpy(z, r, sizeof r);
  }
  void to_mont(uint32_t* z, const Big& x) const {
    uint32_t t[N];
    from_big(t, x);
    mul(z, t, r2);
  }
  Big from_mont(const uint32_t* x) const {
    uint32_t o[N], t[N];
    memset(o, 0, sizeof o);
    o[0] = 1;
    mul(t, x, o);
    reduce(t);
    return to_big(t);
  }
};

// ---------------------------------------------------------------------------------------------
// curve constants handed over by the python side (oracle/port.py), little-endian 64-byte fields
// ---------------------------------------------------------------------------------------------
struct PortParams {
  uint8_t p[64], q[64], beta[64], k2d[64];
  uint8_t v00[64], v01[64], v10[64], v11[64], m0[64], m1[64];  // absolute values
  int32_t v00_neg, v01_neg, v10_neg, v11_neg, m0_neg, m1_neg;
  int32_t glv_m, glv_k, glv_max_bits, scalar_bits, is_te, b3;
};

template <int N>
struct Curve {
  Field<N> F;
  PortParams pp;
  uint32_t beta[N], k2d[N], b3m[N];
  Big q, v00, v01, v10, v11, m0, m1;
  void init(const PortParams& P) {
    pp = P;
    F.init(Big::from_le(P.p, 64));
    q = Big::from_le(P.q, 64);
    F.to_mont(beta, Big::from_le(P.beta, 64));
    F.to_mont(k2d, Big::from_le(P.k2d, 64));
    Big b3;
    b3.v[0] = (uint64_t)P.b3;
    F.to_mont(b3m, b3);
    v00 = Big::from_le(P.v00, 64);
    v01 = Big::from_le(P.v01, 64);
    v10 = Big::from_le(P.v10, 64);
    v11 = Big::from_le(P.v11, 64);
    m0 = Big::from_le(P.m0, 64);
    m1 = Big::from_le(P.m1, 64);
  }
  // src/wasm/glv.ts:68-169 -> |s0|, |s1|, flags
  int glv(const Big& s, Big& s0, Big& s1) const {
    Big sh = big_shr(s, pp.glv_k);
    auto roundmul = [&](const Big& m) {
      Big pr = big_mul(m, sh);
      Big x = big_shr(pr, pp.glv_m);
      if (pr.bit(pp.glv_m - 1)) {
        Big o;
        o.v[0] = 1;
        x = big_add(x, o);
      }
      return x;
    };
    Big x0 = roundmul(m0), x1 = roundmul(m1);  // magnitudes; sign(x_i) = sign(m_i)
    auto term = [&](Big acc, const Big& v, int vneg, const Big& x, int xneg) {
      Big t = big_mul(v, x);
      return (vneg ^ xneg) ? big_sub(acc, t) : big_add(acc, t);
    };
    Big a0 = term(term(s, v00, pp.v00_neg, x0, pp.m0_neg), v01, pp.v01_neg, x1, pp.m1_neg);
    Big a1 = term(term(Big(), v10, pp.v10_neg, x0, pp.m0_neg), v11, pp.v11_neg, x1, pp.m1_neg);
    int flags = 0;
    if (big_neg_flag(a0)) {
      a0 = big_negate(a0);
      flags |= 1;
    }
    if (big_neg_flag(a1)) {
      a1 = big_negate(a1);
      flags |= 2;
    }
    s0 = a0;
    s1 = a1;
    return flags;
  }
};

static inline uint32_t bits_at(const Big& s, int start, int len) {
  Big t = big_shr(s, start);
  return (uint32_t)(t.v[0] & ((1ull << len) - 1));
}

static void parallel_for(int T, const std::function<void(int)>& f) {
  if (T <= 1) {
    f(0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(f, t);
  f(0);
  for (auto& x : th) x.join();
}
// static split of [0, n) like range(), src/threads/threads.ts:354-359
static inline void range_of(size_t n, int t, int T, size_t& lo, size_t& hi) {
  size_t per = (n + T - 1) / T;
  lo = std::min(n, per * t);
  hi = std::min(n, lo + per);
}

// ---------------------------------------------------------------------------------------------
// affine Weierstrass points: x[N] | y[N] | flag   (src/curve-affine.ts:20-52)
// ---------------------------------------------------------------------------------------------
template <int N>
struct Weier {
  typedef Curve<N> C;
  static const int SA = 2 * N + 1;  // words per affine point
  static const int SP = 3 * N + 1;  // words per projective point

  // ---- projective (src/curve-projective.ts) ----
  static void p_zero(uint32_t* P) { memset(P, 0, SP * 4); }
  static bool p_is_zero(const uint32_t* P) { return P[3 * N] == 0; }
  static void p_from_affine(const C& c, uint32_t* P, const uint32_t* A) {
    memcpy(P, A, 2 * N * 4);
    memcpy(P + 2 * N, c.F.one, N * 4);
    P[3 * N] = A[2 * N];
  }
  // dbl-1998-cmo-2, a = 0 (:202-253)
  static void p_double(const C& c, uint32_t* R, const uint32_t* P) {
    const Field<N>& F = c.F;
    if (p_is_zero(P)) {
      p_zero(R);
      return;
    }
    const uint32_t *X = P, *Y = P + N, *Z = P + 2 * N;
    uint32_t w[N], s[N], ss[N], sss[N], Rr[N], B[N], h[N], t[N], u[N];
    F.sqr(t, X);
    F.add(w, t, t);
    F.add(w, w, t);  // 3 X^2
    F.mul(s, Y, Z);
    F.sqr(ss, s);
    F.mul(sss, s, ss);
    F.mul(Rr, Y, s);
    F.mul(B, X, Rr);
    F.sqr(h, w);
    F.add(t, B, B);
    F.add(t, t, t);  // 4B
    F.add(u, t, t);  // 8B
    F.sub(h, h, u);
    uint32_t X3[N], Y3[N], Z3[N];
    F.mul(X3, h, s);
    F.add(X3, X3, X3);
    F.sub(t, t, h);  // 4B - h
    F.mul(Y3, w, t);
    F.sqr(u, Rr);
    F.add(u, u, u);
    F.add(u, u, u);
    F.add(u, u, u);  // 8 R^2
    F.sub(Y3, Y3, u);
    F.add(Z3, sss, sss);
    F.add(Z3, Z3, Z3);
    F.add(Z3, Z3, Z3);
    memcpy(R, X3, N * 4);
    memcpy(R + N, Y3, N * 4);
    memcpy(R + 2 * N, Z3, N * 4);
    R[3 * N] = 1;
  }
  // add-1998-cmo-2 with zero / doubling / inverse handling (:51-160)
  static void p_add(const C& c, uint32_t* R, const uint32_t* P, const uint32_t* Q) {
    const Field<N>& F = c.F;
    if (p_is_zero(P)) {
      memcpy(R, Q, SP * 4);
      return;
    }
    if (p_is_zero(Q)) {
      memcpy(R, P, SP * 4);
      return;
    }
    const uint32_t *X1 = P, *Y1 = P + N, *Z1 = P + 2 * N, *X2 = Q, *Y2 = Q + N, *Z2 = Q + 2 * N;
    uint32_t Y1Z2[N], X1Z2[N], Z1Z2[N], u[N], uu[N], v[N], vv[N], vvv[N], Rr[N], A[N], t[N];
    F.mul(Y1Z2, Y1, Z2);
    F.mul(X1Z2, X1, Z2);
    F.mul(Z1Z2, Z1, Z2);
    F.mul(t, Y2, Z1);
    F.sub(u, t, Y1Z2);
    F.mul(t, X2, Z1);
    F.sub(v, t, X1Z2);
    if (F.is_zero(v)) {
      if (F.is_zero(u))
        p_double(c, R, P);
      else
        p_zero(R);
      return;
    }
    F.sqr(uu, u);
    F.sqr(vv, v);
    F.mul(vvv, v, vv);
    F.mul(Rr, vv, X1Z2);
    F.mul(A, uu, Z1Z2);
    F.sub(A, A, vvv);
    F.sub(A, A, Rr);
    F.sub(A, A, Rr);
    uint32_t X3[N], Y3[N], Z3[N];
    F.mul(X3, v, A);
    F.sub(t, Rr, A);
    F.mul(Y3, u, t);
    F.mul(t, vvv, Y1Z2);
    F.sub(Y3, Y3, t);
    F.mul(Z3, vvv, Z1Z2);
    memcpy(R, X3, N * 4);
    memcpy(R + N, Y3, N * 4);
    memcpy(R + 2 * N, Z3, N * 4);
    R[3 * N] = 1;
  }
  static void p_add_affine(const C& c, uint32_t* R, const uint32_t* P, const uint32_t* A, bool neg) {
    uint32_t Q[SP];
    p_from_affine(c, Q, A);
    if (neg && Q[3 * N]) c.F.sub(Q + N, c.F.p, Q + N);
    p_add(c, R, P, Q);
  }

  // ---- affine batch addition, safe semantics (src/curve-affine.ts:376-458; trick :463-522) ----
  // S[i] = G[i] + H[i] for n independent pairs; pointers into point arrays (SA words each).
  static void batch_add(const C& c, uint32_t** S, uint32_t** G, uint32_t** H, size_t n, std::vector<uint32_t>& scratch) {
    const Field<N>& F = c.F;
    if (n == 0) return;
    // denominators and case per pair
    scratch.resize(n * (N + 1) + 4 * N);
    uint32_t* pre = scratch.data();           // prefix products
    uint32_t* kind = scratch.data() + n * N;  // 0 add, 1 double, 2 take G, 3 take H, 4 zero
    uint32_t run[N], d[N], t[N];
    memcpy(run, F.one, sizeof run);
    for (size_t i = 0; i < n; i++) {
      const uint32_t *g = G[i], *h = H[i];
      uint32_t k;
      if (!h[2 * N]) k = 2;
      else if (!g[2 * N]) k = 3;
      else {
        F.sub(d, h, g);
        if (F.is_zero(d)) {
          if (F.is_equal(g + N, h + N) && !F.is_zero(g + N)) {
            F.add(d, g + N, g + N);
            k = 1;
          } else
            k = 4;
        } else
          k = 0;
      }
      kind[i] = k;
      memcpy(pre + i * N, run, N * 4);
      if (k <= 1) F.mul(run, run, d);
    }
    uint32_t inv[N];
    F.inverse(inv, run);
    for (size_t i = n; i-- > 0;) {
      const uint32_t *g = G[i], *h = H[i];
      uint32_t* s = S[i];
      uint32_t k = kind[i];
      if (k == 2) {
        if (s != g) memcpy(s, g, SA * 4);
        continue;
      }
      if (k == 3) {
        memcpy(s, h, SA * 4);
        continue;
      }
      if (k == 4) {
        memset(s, 0, SA * 4);
        continue;
      }
      uint32_t num[N], m[N], id[N], x3[N], y3[N];
      if (k == 1) {
        F.add(d, g + N, g + N);
        F.sqr(t, g);
        F.add(num, t, t);
        F.add(num, num, t);
      } else {
        F.sub(d, h, g);
        F.sub(num, h + N, g + N);
      }
      F.mul(id, inv, pre + i * N);
      F.mul(inv, inv, d);
      F.mul(m, num, id);
      // src/wasm/curve.ts:32-58 addAffine
      F.sqr(x3, m);
      F.sub(x3, x3, g);
      F.sub(x3, x3, h);  // doubling: h.x == g.x
      F.sub(t, g, x3);
      F.mul(y3, m, t);
      F.sub(y3, y3, g + N);
      memcpy(s, x3, N * 4);
      memcpy(s + N, y3, N * 4);
      s[2 * N] = 1;
    }
  }

  // ---- the MSM (src/msm-batched-affine.ts:74-328) ----
  // points: n affine points (Montgomery limb29, SA words each); scalars: n x 32 bytes LE.
  static void msm_affine(const C& c, const uint8_t* scalars, const uint32_t* points, size_t n, int T, int cw,
                         uint32_t* result /* SP words */) {
    const Field<N>& F = c.F;
    const bool verbose = getenv("MSM_PORT_VERBOSE") != nullptr;
    auto tic = std::chrono::steady_clock::now();
    auto toc = [&](const char* what) {  // the reference's tic/toc log, src/msm-common.ts:192-230
      auto now = std::chrono::steady_clock::now();
      if (verbose) fprintf(stderr, "%-32s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - tic).count());
      tic = now;
    };
    const int b = c.pp.glv_max_bits;
    const int K = (b + 1 + cw - 1) / cw;
    const size_t L = (size_t)1 << (cw - 1);
    const size_t S2 = 2 * n;
    // prepare points & scalars (:338-409): 4 variants per point with sign folding
    std::vector<uint32_t> prep(4 * n * SA);
    std::vector<Big> half(S2);
    parallel_for(T, [&](int t) {
      size_t lo, hi;
      range_of(n, t, T, lo, hi);
      for (size_t i = lo; i < hi; i++) {
        Big s = Big::from_le(scalars + 32 * i, 32);
        while (big_cmp(s, c.q) >= 0) s = big_sub(s, c.q);
        Big s0, s1;
        int flags = c.glv(s, s0, s1);
        half[2 * i] = s0;
        half[2 * i + 1] = s1;
        const uint32_t* P = points + i * SA;
        uint32_t* o = prep.data() + 4 * i * SA;
        uint32_t negy[N], ex[N];
        F.sub(negy, F.p, P + N);
        F.mul(ex, P, c.beta);  // endomorphism, src/wasm/curve.ts:90-103
        const uint32_t* y0 = (flags & 1) ? negy : P + N;
        const uint32_t* y0n = (flags & 1) ? P + N : negy;
        const uint32_t* y1 = (flags & 2) ? negy : P + N;
        const uint32_t* y1n = (flags & 2) ? P + N : negy;
        const uint32_t* xs[4] = {P, P, ex, ex};
        const uint32_t* ys[4] = {y0, y0n, y1, y1n};
        for (int v = 0; v < 4; v++) {
          memcpy(o + v * SA, xs[v], N * 4);
          memcpy(o + v * SA + N, ys[v], N * 4);
          o[v * SA + 2 * N] = P[2 * N];
        }
      }
    });
    toc("prepare points & scalars");
    // slice scalars & count buckets (:166-202)
    std::vector<uint32_t> slices((size_t)K * S2);
    std::vector<std::atomic<uint32_t>> counts((size_t)K * (L + 1));
    for (auto& x : counts) x.store(0, std::memory_order_relaxed);
    parallel_for(T, [&](int t) {
      size_t lo, hi;
      range_of(S2, t, T, lo, hi);
      for (size_t i = lo; i < hi; i++) {
        uint32_t carry = 0;
        for (int k = 0; k < K; k++) {
          uint32_t l = bits_at(half[i], k * cw, cw) + carry;
          if (l > L) {
            l = (uint32_t)(2 * L) - l;
            carry = 1;
          } else
            carry = 0;
          slices[(size_t)k * S2 + i] = l | (carry << 31);
          if (l) counts[(size_t)k * (L + 1) + l].fetch_add(1, std::memory_order_relaxed);
        }
      }
    });
    toc("slice scalars & count buckets");
    // integrate bucket counts (:411-435)
    std::vector<size_t> start((size_t)K * (L + 2));
    size_t maxb = 0;
    for (int k = 0; k < K; k++) {
      size_t run = 0;
      for (size_t l = 1; l <= L; l++) {
        size_t cnt = counts[(size_t)k * (L + 1) + l].load();
        maxb = std::max(maxb, cnt);
        start[(size_t)k * (L + 2) + l] = run;
        run += cnt;
      }
      start[(size_t)k * (L + 2) + L + 1] = run;
    }
    toc("integrate bucket counts");
    // sort points (:444-490): copies the points into bucket order, one array per window
    std::vector<std::vector<uint32_t>> sorted(K);
    parallel_for(std::min(T, K), [&](int t) {
      size_t klo, khi;
      range_of(K, t, std::min(T, K), klo, khi);
      for (size_t k = klo; k < khi; k++) {
        size_t total = start[k * (L + 2) + L + 1];
        sorted[k].resize((total + 1) * SA);
        std::vector<size_t> pos(L + 2);
        for (size_t l = 1; l <= L; l++) pos[l] = start[k * (L + 2) + l];
        for (size_t i = 0; i < S2; i++) {
          uint32_t l = slices[k * S2 + i];
          uint32_t carry = l >> 31;
          l &= 0x7fffffffu;
          if (!l) continue;
          const uint32_t* src = prep.data() + (2 * i + carry) * SA;
          memcpy(sorted[k].data() + pos[l]++ * SA, src, SA * 4);
        }
      }
    });
    toc("sort points");
    // chunks: every window's buckets split into T pieces balanced by point count
    struct Chunk {
      int k;
      size_t l0, l1;
    };
    std::vector<std::vector<Chunk>> chunks(T);
    for (int k = 0; k < K; k++) {
      size_t total = start[(size_t)k * (L + 2) + L + 1];
      size_t l = 1;
      for (int t = 0; t < T; t++) {
        size_t target = total * (t + 1) / T;
        size_t l1 = l;
        while (l1 <= L && start[(size_t)k * (L + 2) + l1 + 1] <= target) l1++;
        if (t == T - 1) l1 = L + 1;
        if (l1 > l) chunks[t].push_back({k, l, l1});
        l = l1;
      }
    }
    // bucket accumulation (:226-271) + local reduction (:284-288, 544-571)
    std::vector<std::vector<uint32_t>> cols(T);
    parallel_for(T, [&](int t) {
      std::vector<uint32_t*> G, H;
      std::vector<uint32_t> scratch;
      for (size_t m = 1; m < maxb; m *= 2) {
        G.clear();
        H.clear();
        for (auto& ch : chunks[t]) {
          uint32_t* base = sorted[ch.k].data();
          const size_t* st = &start[(size_t)ch.k * (L + 2)];
          for (size_t l = ch.l0; l < ch.l1; l++) {
            size_t a = st[l], e = st[l + 1];
            for (size_t i = a; i + m < e; i += 2 * m) {
              G.push_back(base + i * SA);
              H.push_back(base + (i + m) * SA);
            }
          }
        }
        batch_add(c, G.data(), G.data(), H.data(), G.size(), scratch);
      }
      // running sums per chunk: column = triangle + (lstart - 1) * row
      cols[t].assign(chunks[t].size() * SP, 0);
      for (size_t ci = 0; ci < chunks[t].size(); ci++) {
        auto& ch = chunks[t][ci];
        uint32_t* base = sorted[ch.k].data();
        const size_t* st = &start[(size_t)ch.k * (L + 2)];
        uint32_t row[SP], tri[SP];
        p_zero(row);
        p_zero(tri);
        for (size_t l = ch.l1; l-- > ch.l0;) {
          if (st[l + 1] > st[l]) p_add_affine(c, row, row, base + st[l] * SA, false);
          p_add(c, tri, tri, row);
        }
        size_t mult = ch.l0 - 1;
        uint32_t r2[SP];
        memcpy(r2, row, sizeof r2);
        while (mult) {
          if (mult & 1) p_add(c, tri, tri, r2);
          mult >>= 1;
          if (mult) p_double(c, r2, r2);
        }
        memcpy(cols[t].data() + ci * SP, tri, SP * 4);
      }
    });
    toc("bucket accumulation + reduction");
    // partition sums and final Horner (:299-322)
    std::vector<uint32_t> part((size_t)K * SP, 0);
    for (int t = 0; t < T; t++)
      for (size_t ci = 0; ci < chunks[t].size(); ci++)
        p_add(c, part.data() + (size_t)chunks[t][ci].k * SP, part.data() + (size_t)chunks[t][ci].k * SP,
              cols[t].data() + ci * SP);
    uint32_t acc[SP];
    memcpy(acc, part.data() + (size_t)(K - 1) * SP, SP * 4);
    for (int k = K - 2; k >= 0; k--) {
      for (int j = 0; j < cw; j++) p_double(c, acc, acc);
      p_add(c, acc, acc, part.data() + (size_t)k * SP);
    }
    memcpy(result, acc, SP * 4);
    toc("partition sum + final sum");
  }

  // generic bucket method on projective points (msmProjective, src/msm-basic.ts:45-176)
  static void msm_projective(const C& c, const uint8_t* scalars, const uint32_t* points, size_t n, int T, int cw,
                             uint32_t* result) {
    const int b = c.pp.scalar_bits;
    const int K = (b + 1 + cw - 1) / cw;
    const size_t L = (size_t)1 << (cw - 1);
    std::vector<uint32_t> slices((size_t)K * n);
    parallel_for(T, [&](int t) {
      size_t lo, hi;
      range_of(n, t, T, lo, hi);
      for (size_t i = lo; i < hi; i++) {
        Big s = Big::from_le(scalars + 32 * i, 32);
        while (big_cmp(s, c.q) >= 0) s = big_sub(s, c.q);
        uint32_t carry = 0;
        for (int k = 0; k < K; k++) {
          uint32_t l = bits_at(s, k * cw, cw) + carry;
          if (l > L) {
            l = (uint32_t)(2 * L) - l;
            carry = 1;
          } else
            carry = 0;
          slices[(size_t)k * n + i] = l | (carry << 31);
        }
      }
    });
    std::vector<uint32_t> cols((size_t)T * K * SP, 0);
    parallel_for(T, [&](int t) {
      for (int k = 0; k < K; k++) {
        size_t lo, hi;
        range_of(L, t, T, lo, hi);  // buckets lo+1 .. hi
        if (hi <= lo) continue;
        std::vector<uint32_t> buckets((hi - lo) * SP, 0);
        for (size_t i = 0; i < n; i++) {
          uint32_t l = slices[(size_t)k * n + i];
          uint32_t carry = l >> 31;
          l &= 0x7fffffffu;
          if (l <= lo || l > hi) continue;
          uint32_t* B = buckets.data() + (l - lo - 1) * SP;
          p_add_affine(c, B, B, points + i * SA, carry);
        }
        uint32_t row[SP], tri[SP];
        p_zero(row);
        p_zero(tri);
        for (size_t l = hi; l-- > lo;) {
          p_add(c, row, row, buckets.data() + (l - lo) * SP);
          p_add(c, tri, tri, row);
        }
        size_t mult = lo;
        uint32_t r2[SP];
        memcpy(r2, row, sizeof r2);
        while (mult) {
          if (mult & 1) p_add(c, tri, tri, r2);
          mult >>= 1;
          if (mult) p_double(c, r2, r2);
        }
        memcpy(cols.data() + ((size_t)t * K + k) * SP, tri, SP * 4);
      }
    });
    std::vector<uint32_t> part((size_t)K * SP, 0);
    for (int t = 0; t < T; t++)
      for (int k = 0; k < K; k++) p_add(c, part.data() + (size_t)k * SP, part.data() + (size_t)k * SP, cols.data() + ((size_t)t * K + k) * SP);
    uint32_t acc[SP];
    memcpy(acc, part.data() + (size_t)(K - 1) * SP, SP * 4);
    for (int k = K - 2; k >= 0; k--) {
      for (int j = 0; j < cw; j++) p_double(c, acc, acc);
      p_add(c, acc, acc, part.data() + (size_t)k * SP);
    }
    memcpy(result, acc, SP * 4);
  }
};

// ---------------------------------------------------------------------------------------------
// twisted Edwards, extended coordinates X|Y|Z|T (src/curve-twisted-edwards.ts:84-165)
// ---------------------------------------------------------------------------------------------
template <int N>
struct Te {
  typedef Curve<N> C;
  static const int SE = 4 * N;
  static void zero(const C& c, uint32_t* P) {
    memset(P, 0, SE * 4);
    memcpy(P + N, c.F.one, N * 4);
    memcpy(P + 2 * N, c.F.one, N * 4);
  }
  // unified add; `mixed`: Z2 = 1; `neg`: subtract
  static void add(const C& c, uint32_t* R, const uint32_t* P, const uint32_t* Q, bool mixed, bool neg) {
    const Field<N>& F = c.F;
    const uint32_t *X1 = P, *Y1 = P + N, *Z1 = P + 2 * N, *T1 = P + 3 * N;
    const uint32_t *X2 = Q, *Y2 = Q + N, *Z2 = Q + 2 * N, *T2 = Q + 3 * N;
    uint32_t A[N], B[N], Cc[N], D[N], E[N], Fv[N], G[N], H[N], t[N], u[N];
    F.sub(t, Y1, X1);
    F.add(u, Y1, X1);
    if (!neg) {
      uint32_t a2[N], b2[N];
      F.sub(a2, Y2, X2);
      F.add(b2, Y2, X2);
      F.mul(A, t, a2);
      F.mul(B, u, b2);
    } else {
      uint32_t a2[N], b2[N];
      F.add(a2, Y2, X2);
      F.sub(b2, Y2, X2);
      F.mul(A, t, a2);
      F.mul(B, u, b2);
    }
    F.mul(Cc, T1, T2);
    F.mul(Cc, Cc, c.k2d);
    if (neg) {
      uint32_t z[N];
      memset(z, 0, sizeof z);
      F.sub(Cc, z, Cc);
    }
    if (mixed)
      F.add(D, Z1, Z1);
    else {
      F.mul(D, Z1, Z2);
      F.add(D, D, D);
    }
    F.sub(E, B, A);
    F.sub(Fv, D, Cc);
    F.add(G, D, Cc);
    F.add(H, B, A);
    uint32_t X3[N], Y3[N], Z3[N], T3[N];
    F.mul(X3, E, Fv);
    F.mul(Y3, G, H);
    F.mul(T3, E, H);
    F.mul(Z3, Fv, G);
    memcpy(R, X3, N * 4);
    memcpy(R + N, Y3, N * 4);
    memcpy(R + 2 * N, Z3, N * 4);
    memcpy(R + 3 * N, T3, N * 4);
  }
  // src/msm-basic.ts:45-176
  static void msm_basic(const C& c, const uint8_t* scalars, const uint32_t* points, size_t n, int T, int cw, uint32_t* result) {
    const int b = c.pp.scalar_bits;
    const int K = (b + 1 + cw - 1) / cw;
    const size_t L = (size_t)1 << (cw - 1);
    std::vector<uint32_t> slices((size_t)K * n);
    parallel_for(T, [&](int t) {
      size_t lo, hi;
      range_of(n, t, T, lo, hi);
      for (size_t i = lo; i < hi; i++) {
        Big s = Big::from_le(scalars + 32 * i, 32);
        while (big_cmp(s, c.q) >= 0) s = big_sub(s, c.q);
        uint32_t carry = 0;
        for (int k = 0; k < K; k++) {
          uint32_t l = bits_at(s, k * cw, cw) + carry;
          if (l > L) {
            l = (uint32_t)(2 * L) - l;
            carry = 1;
          } else
            carry = 0;
          slices[(size_t)k * n + i] = l | (carry << 31);
        }
      }
    });
    std::vector<uint32_t> cols((size_t)T * K * SE);
    parallel_for(T, [&](int t) {
      for (int k = 0; k < K; k++) {
        uint32_t* col = cols.data() + ((size_t)t * K + k) * SE;
        zero(c, col);
        size_t lo, hi;
        range_of(L, t, T, lo, hi);
        if (hi <= lo) continue;
        std::vector<uint32_t> buckets((hi - lo) * SE);
        for (size_t j = 0; j < hi - lo; j++) zero(c, buckets.data() + j * SE);
        for (size_t i = 0; i < n; i++) {
          uint32_t l = slices[(size_t)k * n + i];
          uint32_t carry = l >> 31;
          l &= 0x7fffffffu;
          if (l <= lo || l > hi) continue;
          uint32_t* B = buckets.data() + (l - lo - 1) * SE;
          add(c, B, B, points + i * SE, true, carry);
        }
        uint32_t row[SE], tri[SE];
        zero(c, row);
        zero(c, tri);
        for (size_t l = hi; l-- > lo;) {
          add(c, row, row, buckets.data() + (l - lo) * SE, false, false);
          add(c, tri, tri, row, false, false);
        }
        size_t mult = lo;
        uint32_t r2[SE];
        memcpy(r2, row, sizeof r2);
        while (mult) {
          if (mult & 1) add(c, tri, tri, r2, false, false);
          mult >>= 1;
          if (mult) add(c, r2, r2, r2, false, false);
        }
        memcpy(col, tri, SE * 4);
      }
    });
    std::vector<uint32_t> part((size_t)K * SE);
    for (int k = 0; k < K; k++) zero(c, part.data() + (size_t)k * SE);
    for (int t = 0; t < T; t++)
      for (int k = 0; k < K; k++) add(c, part.data() + (size_t)k * SE, part.data() + (size_t)k * SE, cols.data() + ((size_t)t * K + k) * SE, false, false);
    uint32_t acc[SE];
    memcpy(acc, part.data() + (size_t)(K - 1) * SE, SE * 4);
    for (int k = K - 2; k >= 0; k--) {
      for (int j = 0; j < cw; j++) add(c, acc, acc, acc, false, false);
      add(c, acc, acc, part.data() + (size_t)k * SE, false, false);
    }
    memcpy(result, acc, SE * 4);
  }
};


// ---------------------------------------------------------------------------------------------
// Seeded synthetic inputs on the CPU: the same counter-based construction as the CUDA generators
// (msm_zprize_b200/csrc/kernels_basic.cuh k_rp_tables / k_rp_points / k_random_scalars, which replace
// randomPointsFast / randomScalars, src/curve-random.ts:24-91,151-194), so that the reference arm of
// bench.py gets byte-identical inputs without loading the CUDA library.
// ---------------------------------------------------------------------------------------------
static inline uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static const int RP_TABLES = 4, RP_BITS = 13;

template <int N>
static void gen_scalars(const Curve<N>& c, uint8_t* out, size_t first, size_t n, uint64_t seed, int T) {
  const int qbits = c.pp.scalar_bits;
  parallel_for(T, [&](int t) {
    size_t lo, hi;
    range_of(n, t, T, lo, hi);
    for (size_t i = lo; i < hi; i++) {
      uint64_t st = seed + (uint64_t)(first + i + 1) * 0xD1342543DE82EF95ull;
      Big s;
      for (;;) {
        s = Big();
        for (int j = 0; j < 4; j++) s.v[j] = splitmix64(st);
        const int top = qbits - 224;  // bits kept in the top 32-bit word
        uint64_t m = top >= 32 ? 0xFFFFFFFFull : ((1ull << top) - 1);
        s.v[3] = (s.v[3] & 0xFFFFFFFFull) | ((s.v[3] >> 32) & m) << 32;
        if (big_cmp(s, c.q) < 0) break;
      }
      s.to_le(out + 32 * i, 32);
    }
  });
}

// point arithmetic the generator needs, for both curve forms: accumulators are projective (Weierstrass,
// 3N+1 words) or extended (twisted Edwards, 4N words); table entries are affine x | y (2N words, Montgomery)
template <int N>
struct GenOps {
  typedef Curve<N> C;
  static int acc_words(const C& c) { return c.pp.is_te ? 4 * N : 3 * N + 1; }
  static void zero(const C& c, uint32_t* A) {
    if (c.pp.is_te) Te<N>::zero(c, A);
    else Weier<N>::p_zero(A);
  }
  static void add_xy(const C& c, uint32_t* A, const uint32_t* xy) {
    if (c.pp.is_te) {
      uint32_t Q[4 * N];
      memcpy(Q, xy, 2 * N * 4);
      memcpy(Q + 2 * N, c.F.one, N * 4);
      c.F.mul(Q + 3 * N, xy, xy + N);
      Te<N>::add(c, A, A, Q, true, false);
    } else {
      uint32_t Q[2 * N + 1];
      memcpy(Q, xy, 2 * N * 4);
      Q[2 * N] = 1;
      Weier<N>::p_add_affine(c, A, A, Q, false);
    }
  }
  static void dbl(const C& c, uint32_t* A) {
    if (c.pp.is_te) Te<N>::add(c, A, A, A, false, false);
    else Weier<N>::p_double(c, A, A);
  }
  static bool is_inf(const C& c, const uint32_t* A) { return !c.pp.is_te && (A[3 * N] == 0 || c.F.is_zero(A + 2 * N)); }
  // affine x | y (Montgomery, reduced) of m accumulators with one shared inversion
  static void normalise(const C& c, const uint32_t* A, size_t m, uint32_t* xy) {
    const Field<N>& F = c.F;
    const int AW = acc_words(c);
    std::vector<uint32_t> pre(m * N);
    uint32_t run[N], inv[N], zi[N];
    memcpy(run, F.one, sizeof run);
    for (size_t u = 0; u < m; u++) {
      memcpy(pre.data() + u * N, run, N * 4);
      F.mul(run, run, A + u * AW + 2 * N);
    }
    F.inverse(inv, run);
    for (size_t u = m; u-- > 0;) {
      F.mul(zi, inv, pre.data() + u * N);
      F.mul(inv, inv, A + u * AW + 2 * N);
      F.mul(xy + u * 2 * N, A + u * AW, zi);
      F.mul(xy + u * 2 * N + N, A + u * AW + N, zi);
      F.reduce(xy + u * 2 * N);
      F.reduce(xy + u * 2 * N + N);
    }
  }
};

template <int N>
static void gen_points(const Curve<N>& c, const uint8_t* gx_le, const uint8_t* gy_le, uint8_t* out, size_t first, size_t n,
                       uint64_t seed, int nb, int T) {
  typedef GenOps<N> G;
  const Field<N>& F = c.F;
  const int AW = G::acc_words(c);
  uint32_t gxy[2 * N];
  F.to_mont(gxy, Big::from_le(gx_le, 64));
  F.to_mont(gxy + N, Big::from_le(gy_le, 64));
  F.reduce(gxy);
  F.reduce(gxy + N);
  // tables: entry (k, j) = (j + 1) * B_k, B_k = h_k * G
  const size_t E = (size_t)1 << RP_BITS;
  std::vector<uint32_t> tables((size_t)RP_TABLES * E * 2 * N);
  parallel_for(std::min(T, RP_TABLES), [&](int t) {
    size_t klo, khi;
    range_of(RP_TABLES, t, std::min(T, RP_TABLES), klo, khi);
    for (size_t k = klo; k < khi; k++) {
      uint64_t st = seed ^ (0xA5A5A5A5ull + k);
      uint64_t h = splitmix64(st) | 1ull;
      std::vector<uint32_t> acc(AW), bk(2 * N);
      G::zero(c, acc.data());
      bool started = false;
      for (int bit = 63; bit >= 0; bit--) {
        if (started) G::dbl(c, acc.data());
        if ((h >> bit) & 1) {
          G::add_xy(c, acc.data(), gxy);
          started = true;
        }
      }
      G::normalise(c, acc.data(), 1, bk.data());
      // multiples of B_k, normalised in batches
      const size_t BATCH = 256;
      std::vector<uint32_t> accs(BATCH * AW);
      std::vector<uint32_t> cur(AW);
      G::zero(c, cur.data());
      for (size_t j0 = 0; j0 < E; j0 += BATCH) {
        for (size_t u = 0; u < BATCH; u++) {
          G::add_xy(c, cur.data(), bk.data());
          memcpy(accs.data() + u * AW, cur.data(), AW * 4);
        }
        G::normalise(c, accs.data(), BATCH, tables.data() + (k * E + j0) * 2 * N);
      }
    }
  });
  const uint64_t pseed = seed ^ 0x5EEDull;
  parallel_for(T, [&](int t) {
    size_t lo, hi;
    range_of(n, t, T, lo, hi);
    const size_t BATCH = 256;
    std::vector<uint32_t> accs(BATCH * AW), xy(BATCH * 2 * N);
    for (size_t i0 = lo; i0 < hi; i0 += BATCH) {
      size_t m = std::min(BATCH, hi - i0);
      for (size_t u = 0; u < m; u++) {
        uint64_t st = pseed + (uint64_t)(first + i0 + u + 1) * 0xD1342543DE82EF95ull;
        uint64_t r = splitmix64(st);
        uint32_t* acc = accs.data() + u * AW;
        G::zero(c, acc);
        for (int k = 0; k < RP_TABLES; k++) {
          uint32_t w = (uint32_t)(r >> (RP_BITS * k)) & ((1u << RP_BITS) - 1u);
          G::add_xy(c, acc, tables.data() + ((size_t)k * E + w) * 2 * N);
        }
        if (G::is_inf(c, acc)) {
          G::zero(c, acc);
          G::add_xy(c, acc, gxy);
        }
      }
      G::normalise(c, accs.data(), m, xy.data());
      for (size_t u = 0; u < m; u++) {
        F.from_mont(xy.data() + u * 2 * N).to_le(out + (i0 + u) * 2 * nb, nb);
        F.from_mont(xy.data() + u * 2 * N + N).to_le(out + (i0 + u) * 2 * nb + nb, nb);
      }
    }
  });
}

// ---------------------------------------------------------------------------------------------
// C API (loaded with ctypes by oracle/port.py)
// ---------------------------------------------------------------------------------------------
struct PortCtx {
  int n29;
  Curve<14> c14;
  Curve<9> c9;
};

template <int N>
static void points_from_le(const Curve<N>& c, const uint8_t* le, size_t n, int nb, uint32_t* out, int T) {
  parallel_for(T, [&](int t) {
    size_t lo, hi;
    range_of(n, t, T, lo, hi);
    for (size_t i = lo; i < hi; i++) {
      if (c.pp.is_te) {
        uint32_t* o = out + i * 4 * N;
        c.F.to_mont(o, Big::from_le(le + i * 2 * nb, nb));
        c.F.to_mont(o + N, Big::from_le(le + i * 2 * nb + nb, nb));
        memcpy(o + 2 * N, c.F.one, N * 4);
        c.F.mul(o + 3 * N, o, o + N);
      } else {
        uint32_t* o = out + i * (2 * N + 1);
        c.F.to_mont(o, Big::from_le(le + i * 2 * nb, nb));
        c.F.to_mont(o + N, Big::from_le(le + i * 2 * nb + nb, nb));
        o[2 * N] = 1;
      }
    }
  });
}

template <int N>
static int run_msm(const Curve<N>& c, const uint8_t* scalars, const uint32_t* points, size_t n, int T, int cw, int form,
                   uint8_t* out_x, uint8_t* out_y, int nb) {
  const Field<N>& F = c.F;
  if (c.pp.is_te) {
    uint32_t R[4 * N];
    if (n == 0)
      Te<N>::zero(c, R);
    else
      Te<N>::msm_basic(c, scalars, points, n, T, cw, R);
    uint32_t zi[N], x[N], y[N];
    F.inverse(zi, R + 2 * N);
    F.mul(x, R, zi);
    F.mul(y, R + N, zi);
    F.from_mont(x).to_le(out_x, nb);
    F.from_mont(y).to_le(out_y, nb);
    Big bx = F.from_mont(x), by = F.from_mont(y), one;
    one.v[0] = 1;
    return (big_cmp(bx, Big()) == 0 && big_cmp(by, one) == 0) ? 1 : 0;
  }
  uint32_t R[3 * N + 1];
  if (n == 0)
    Weier<N>::p_zero(R);
  else if (form == 0)
    Weier<N>::msm_affine(c, scalars, points, n, T, cw, R);
  else
    Weier<N>::msm_projective(c, scalars, points, n, T, cw, R);
  memset(out_x, 0, nb);
  memset(out_y, 0, nb);
  if (R[3 * N] == 0 || F.is_zero(R + 2 * N)) return 1;
  uint32_t zi[N], x[N], y[N];
  F.inverse(zi, R + 2 * N);
  F.mul(x, R, zi);
  F.mul(y, R + N, zi);
  F.from_mont(x).to_le(out_x, nb);
  F.from_mont(y).to_le(out_y, nb);
  return 0;
}

extern "C" {
void* port_create(const PortParams* pp, int n29) {
  PortCtx* ctx = new PortCtx();
  ctx->n29 = n29;
  if (n29 == 14)
    ctx->c14.init(*pp);
  else
    ctx->c9.init(*pp);
  return ctx;
}
void port_destroy(void* h) { delete (PortCtx*)h; }
size_t port_point_words(void* h) {
  PortCtx* ctx = (PortCtx*)h;
  int N = ctx->n29;
  bool te = (N == 14 ? ctx->c14.pp.is_te : ctx->c9.pp.is_te);
  return te ? 4 * N : 2 * N + 1;
}
// LE bytes -> the reference's in-memory point layout (Montgomery, 29-bit limbs); untimed set-up
void port_points_from_bytes(void* h, const uint8_t* le, size_t n, int nbytes, uint32_t* out, int threads) {
  PortCtx* ctx = (PortCtx*)h;
  if (ctx->n29 == 14)
    points_from_le(ctx->c14, le, n, nbytes, out, threads);
  else
    points_from_le(ctx->c9, le, n, nbytes, out, threads);
}
// the timed call: returns is_zero; *seconds = wall time of the MSM proper
int port_msm(void* h, const uint8_t* scalars_le, const uint32_t* points, size_t n, int threads, int window_bits, int form,
             uint8_t* out_x, uint8_t* out_y, int nbytes, double* seconds) {
  PortCtx* ctx = (PortCtx*)h;
  auto t0 = std::chrono::steady_clock::now();
  int z;
  if (ctx->n29 == 14)
    z = run_msm(ctx->c14, scalars_le, points, n, threads, window_bits, form, out_x, out_y, nbytes);
  else
    z = run_msm(ctx->c9, scalars_le, points, n, threads, window_bits, form, out_x, out_y, nbytes);
  if (seconds) *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return z;
}
// seeded inputs, byte-identical to msm_b200_random_scalars / msm_b200_random_points (LE_BYTES layout)
// (`first`: global index of the first element, so that a range of a larger set can be generated)
void port_random_scalars(void* h, uint8_t* out_le, size_t first, size_t n, uint64_t seed, int threads) {
  PortCtx* ctx = (PortCtx*)h;
  if (ctx->n29 == 14)
    gen_scalars(ctx->c14, out_le, first, n, seed, threads);
  else
    gen_scalars(ctx->c9, out_le, first, n, seed, threads);
}
void port_random_points(void* h, const uint8_t* gx_le64, const uint8_t* gy_le64, uint8_t* out_le, size_t first, size_t n,
                        uint64_t seed, int nbytes, int threads) {
  PortCtx* ctx = (PortCtx*)h;
  if (ctx->n29 == 14)
    gen_points(ctx->c14, gx_le64, gy_le64, out_le, first, n, seed, nbytes, threads);
  else
    gen_points(ctx->c9, gx_le64, gy_le64, out_le, first, n, seed, nbytes, threads);
}
// single field multiplication throughput (ns per Montgomery product), for the record
double port_mul_ns(void* h, int iters) {
  PortCtx* ctx = (PortCtx*)h;
  auto t0 = std::chrono::steady_clock::now();
  if (ctx->n29 == 14) {
    uint32_t x[14], y[14];
    memcpy(x, ctx->c14.F.r2, sizeof x);
    memcpy(y, ctx->c14.F.one, sizeof y);
    for (int i = 0; i < iters; i++) ctx->c14.F.mul(x, x, y), ctx->c14.F.mul(y, y, x);
    if (x[0] == 0xffffffffu) return -1;
  } else {
    uint32_t x[9], y[9];
    memcpy(x, ctx->c9.F.r2, sizeof x);
    memcpy(y, ctx->c9.F.one, sizeof y);
    for (int i = 0; i < iters; i++) ctx->c9.F.mul(x, x, y), ctx->c9.F.mul(y, y, x);
    if (x[0] == 0xffffffffu) return -1;
  }
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e9 / (2.0 * iters);
}
}
