// Kernels of the generic bucket method (no GLV, no batching): the reference's msmBasic,
// src/msm-basic.ts:45-176, used for twisted Edwards (src/parallel.ts:193-199) and for
// Parallel.msmProjective on Weierstrass curves (src/parallel.ts:69-87).
//
//   k_te_ingest        Parallel.pointsFromBytes (TE), src/parallel.ts:209-232, caching (y+x, y-x, 2dxy)
//   k_load_scalars     scalarsFromBytes / limb unpacking, src/parallel.ts:235-249
//   k_hist_scatter8    digit extraction :80-95 + counting sort (the reference rescans all N digits
//                      per bucket chunk, :115-128; here the indices are sorted once)
//   k_bucket_acc       per-bucket accumulation with addMixed / subMixed, :115-128
// plus the seeded input generators replacing randomPointsFast / randomScalars
// (src/curve-random.ts:24-91,151-194).
#pragma once
#include "kernels_reduce.cuh"

namespace msm {

constexpr int BUCKET_SPLIT = 128;  // entries per virtual bucket of the generic bucket method

template <class F>
__global__ void k_te_ingest(const uint8_t* __restrict__ in, size_t n, int layout, uint4* __restrict__ bases) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<F> x, y;
  if (layout == 0) {  // X | Y | Z | T in 29-bit limbs; inputs are affine (Z = R mod p), :141-146
    const uint32_t* w = reinterpret_cast<const uint32_t*>(in) + i * (4 * F::N29);
    x = fe_from_limb29<F>(w);
    y = fe_from_limb29<F>(w + F::N29);
  } else {
    int nb = (F::BITS + 7) / 8;
    const uint8_t* b = in + i * (size_t)(2 * nb);
    x = fe_from_le_bytes<F>(b, nb);
    y = fe_from_le_bytes<F>(b + nb, nb);
  }
  Niels<F> q = niels_from_xy(x, y);
  uint4* o = bases + i * (size_t)(3 * F::N / 4);
  st_aos<F>(o, q.yp);
  st_aos<F>(o + F::N / 4, q.ym);
  st_aos<F>(o + 2 * F::N / 4, q.kt);
}

// Window tables for resident twisted-Edwards bases (see k_build_table in kernels_weierstrass.cuh): table k
// holds 2^(kc) P_i in the cached form (y+x, y-x, 2dxy).  One thread per point: the previous table's point as
// the extended point (4x : 4y : 4 : 4xy) -- no halving needed -- c doublings, one inversion per block.
template <class F>
__global__ void __launch_bounds__(256) k_te_build_table(const uint4* __restrict__ prev, uint4* __restrict__ next, size_t n, int c) {
  __shared__ uint32_t smem[97 * F::N + F::N];
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const bool valid = i < n;
  Ext<F> A = ext_zero<F>();
  if (valid) {
    const uint4* r = prev + i * (size_t)(3 * F::N / 4);
    const Fe<F> yp = ld_aos<F>(r), ym = ld_aos<F>(r + F::N / 4);
    const Fe<F> dx = fe_sub(yp, ym), sy = fe_add(yp, ym);  // 2x, 2y
    A.X = fe_dbl(dx);
    A.Y = fe_dbl(sy);
    A.Z = fe_dbl(fe_dbl(fe_one<F>()));
    A.T = fe_mul(dx, sy);
  }
#pragma unroll 1
  for (int d = 0; d < c; d++) A = ext_dbl<F>(A);
  Fe<F> others, total;
  block_products<F, 256>(A.Z, others, total, smem);  // Z != 0 on a complete Edwards curve
  uint32_t* binv = smem + 97 * F::N;
  __syncthreads();
  if (threadIdx.x == 0) fe_to_smem<F>(binv, fe_inv(total));
  __syncthreads();
  if (!valid) return;
  const Fe<F> zi = fe_mul(others, fe_from_smem<F>(binv));
  Niels<F> q = niels_from_xy(fe_mul(A.X, zi), fe_mul(A.Y, zi));
  uint4* o = next + i * (size_t)(3 * F::N / 4);
  st_aos<F>(o, q.yp);
  st_aos<F>(o + F::N / 4, q.ym);
  st_aos<F>(o + 2 * F::N / 4, q.kt);
}

// scalars -> 8 limbs in [0, q), two uint4 per scalar
template <class S>
__global__ void k_load_scalars(const uint8_t* __restrict__ in, size_t n, int layout, uint4* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s[8];
  load_scalar(in, i, layout, s);
  scalar_reduce<S>(s);
  out[2 * i] = make_uint4(s[0], s[1], s[2], s[3]);
  out[2 * i + 1] = make_uint4(s[4], s[5], s[6], s[7]);
}

template <bool SCATTER>
__global__ void k_hist_scatter8(SortArgs a) {
  size_t h = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= a.S) return;
  uint4 q0 = a.hs[2 * h], q1 = a.hs[2 * h + 1];
  uint32_t s[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
  uint32_t carry = 0;
  for (int k = 0; k < a.K; k++) {
    uint32_t l = signed_digit<8>(s, k, a.c, carry);
    if (l == 0) continue;
    uint32_t b = (uint32_t)k * a.bucket_stride + (l - 1);
    if (!SCATTER) {
      atomicAdd(&a.cnt[b], 1u);
    } else {
      uint32_t pos = atomicAdd(&a.cursor[b], 1u);
      a.ent[2u * a.po0[b] + pos] = ((uint32_t)h + (uint32_t)k * a.ent_stride) | (carry << 31);
    }
  }
}

// One thread per bucket: acc = sum of +/- base points of the bucket.
template <class C>
__global__ void __launch_bounds__(128) k_bucket_acc(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ po0,
                                                    const uint32_t* __restrict__ ent, const uint4* __restrict__ bases,
                                                    uint32_t NB, uint4* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  typename C::Acc acc = C::zero();
  uint32_t n = cnt[b];
  const uint32_t* e = ent + 2 * (size_t)po0[b];
  if constexpr (C::PREFETCH_BASE) {  // fetch the next base point while adding the current one
    if (n) {
      uint32_t en = e[0];
      typename C::Base cur = C::ld_base(bases, ent_index(en));
#pragma unroll 1
      for (uint32_t j = 0; j < n; j++) {
        const uint32_t en_next = (j + 1 < n) ? e[j + 1] : en;
        const typename C::Base nxt = C::ld_base(bases, ent_index(en_next));
        acc = C::add_cached(acc, cur, ent_neg(en));
        cur = nxt;
        en = en_next;
      }
    }
  } else {
#pragma unroll 1
    for (uint32_t j = 0; j < n; j++) {
      uint32_t en = e[j];
      acc = C::add_base(acc, bases, ent_index(en), ent_neg(en));
    }
  }
  C::st(buckets + (size_t)b * (C::ACC_FE * C::F::N / 4), acc);
}

// Load balancing for skewed bucket sizes (short top window, repeated scalars): a bucket with more
// than `split` entries is cut into virtual buckets of `split` entries (one thread each, k_bucket_acc_v),
// whose accumulators are then added up per bucket (k_bucket_combine).
template <class C>
__global__ void __launch_bounds__(128) k_bucket_acc_v(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ po0,
                                                      const uint32_t* __restrict__ voff, const uint32_t* __restrict__ vkey,
                                                      const uint32_t* __restrict__ ent, const uint4* __restrict__ bases,
                                                      uint32_t V, uint32_t split, uint4* __restrict__ vacc) {
  uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  uint32_t b = vkey[v];
  uint32_t lo = (v - voff[b]) * split;
  uint32_t hi = min(cnt[b], lo + split);
  typename C::Acc acc = C::zero();
  const uint32_t* e = ent + 2 * (size_t)po0[b];
  if constexpr (C::PREFETCH_BASE) {
    // the next base point is fetched while the current one is added (random 96-byte gathers)
    if (lo < hi) {
      uint32_t en = e[lo];
      typename C::Base cur = C::ld_base(bases, ent_index(en));
#pragma unroll 1
      for (uint32_t j = lo; j < hi; j++) {
        const uint32_t en_next = (j + 1 < hi) ? e[j + 1] : en;
        const typename C::Base nxt = C::ld_base(bases, ent_index(en_next));
        acc = C::add_cached(acc, cur, ent_neg(en));
        cur = nxt;
        en = en_next;
      }
    }
  } else {
#pragma unroll 1
    for (uint32_t j = lo; j < hi; j++) {
      uint32_t en = e[j];
      acc = C::add_base(acc, bases, ent_index(en), ent_neg(en));
    }
  }
  C::st(vacc + (size_t)v * (C::ACC_FE * C::F::N / 4), acc);
}

// One halving pass over the virtual accumulators of every bucket (they are contiguous):
// piece i of a bucket absorbs piece i + 2^p when i is a multiple of 2^(p+1).  After
// ceil(log2(max pieces)) passes piece 0 holds the bucket sum -- used instead of the serial loop of
// k_bucket_combine when a bucket has many pieces (heavily skewed scalars).
template <class C>
__global__ void __launch_bounds__(128) k_bucket_tree_pass(const uint32_t* __restrict__ voff, const uint32_t* __restrict__ vkey,
                                                          uint4* __restrict__ vacc, uint32_t V, uint32_t Vtot, int p) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;  // candidate piece index v = t << (p + 1) is not bucket aligned,
  if (t >= V) return;                                  // so every piece checks its own position
  uint32_t b = vkey[t];
  uint32_t i = t - voff[b];
  if (i & ((2u << p) - 1u)) return;
  uint32_t other = t + (1u << p);
  if (other >= Vtot || vkey[other] != b) return;
  constexpr int U4 = C::ACC_FE * C::F::N / 4;
  C::st(vacc + (size_t)t * U4, C::add(C::ld(vacc + (size_t)t * U4), C::ld(vacc + (size_t)other * U4)));
}

template <class C>
__global__ void __launch_bounds__(128) k_bucket_combine(const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ voff,
                                                        const uint4* __restrict__ vacc, uint32_t NB, uint32_t split,
                                                        int serial_max, uint4* __restrict__ buckets) {
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  uint32_t nv = (cnt[b] + split - 1) / split;
  if (nv > (uint32_t)serial_max) nv = 1;  // already folded into piece 0 by the tree passes
  constexpr int U4 = C::ACC_FE * C::F::N / 4;
  typename C::Acc acc = C::zero();
  if (nv) acc = C::ld(vacc + (size_t)voff[b] * U4);
#pragma unroll 1
  for (uint32_t s = 1; s < nv; s++) acc = C::add(acc, C::ld(vacc + (size_t)(voff[b] + s) * U4));
  C::st(buckets + (size_t)b * U4, acc);
}

// ------------------------------------------------------------------------------------------
// seeded synthetic inputs
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  s += 0x9E3779B97F4A7C15ull;
  uint64_t z = s;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Uniform scalar in [0, q): 32 random bytes, top bits masked to the bit length of q, rejection
// sampling -- the distribution of src/curve-random.ts:151-194.  Stream i starts at
// seed + (i+1) * 0xD1342543DE82EF95.
template <class S>
__global__ void k_random_scalars(uint32_t* __restrict__ out, size_t n, uint64_t seed, size_t first) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t st = seed + (uint64_t)(first + i + 1) * 0xD1342543DE82EF95ull;  // `first`: global index of out[0]
  uint32_t s[8];
  for (;;) {
    for (int j = 0; j < 4; j++) {
      uint64_t r = splitmix64(st);
      s[2 * j] = (uint32_t)r;
      s[2 * j + 1] = (uint32_t)(r >> 32);
    }
    constexpr int top = S::QBITS - 224;  // bits kept in limb 7
    s[7] &= (top >= 32) ? 0xFFFFFFFFu : ((1u << top) - 1u);
    if (!scalar_geq_q<S>(s)) break;
  }
  for (int j = 0; j < 8; j++) out[i * 8 + j] = s[j];
}

constexpr int RP_TABLES = 4;
constexpr int RP_BITS = 13;

template <class F>
__device__ __forceinline__ Fe<F> fe_gen(int which) {
  Fe<F> g;
#pragma unroll
  for (int i = 0; i < F::N; i++) g.v[i] = which ? F::GY(i) : F::GX(i);
  return g;
}

// acc + (x, y) for both curve forms
template <class C>
struct AddXY;
template <class F, uint32_t B3>
struct AddXY<WeierCurve<F, B3>> {
  __device__ static Proj<F> run(const Proj<F>& a, const Fe<F>& x, const Fe<F>& y) {
    Aff<F> Q;
    Q.x = x;
    Q.y = y;
    return proj_add_mixed<F, B3>(a, Q);
  }
};
template <class F>
struct AddXY<TeCurve<F>> {
  __device__ static Ext<F> run(const Ext<F>& a, const Fe<F>& x, const Fe<F>& y) {
    return ext_add_niels<F>(a, niels_from_xy(x, y), false);
  }
};

// Table entry (k, j) = (j+1) * B_k as affine (x | y), B_k = h_k * G with h_k a seeded 64-bit value.
// The construction follows randomPointsFast (src/curve-random.ts:24-91): every generated point
// is a sum of one multiple per table.
template <class C>
__global__ void k_rp_tables(uint4* __restrict__ tables, uint64_t seed) {
  using F = typename C::F;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= RP_TABLES * (1u << RP_BITS)) return;
  uint32_t k = t >> RP_BITS, j = (t & ((1u << RP_BITS) - 1u)) + 1;
  uint64_t st = seed ^ (0xA5A5A5A5ull + k);
  uint64_t h = splitmix64(st) | 1ull;
  // scalar = h * j  (< 2^78)
  unsigned __int128 sc = (unsigned __int128)h * j;
  Fe<F> gx = fe_gen<F>(0), gy = fe_gen<F>(1);
  typename C::Acc acc = C::zero();
  bool started = false;
#pragma unroll 1
  for (int bit = 79; bit >= 0; bit--) {
    if (started) acc = C::canon(C::dbl(acc));  // the generator works on canonical values throughout
    if ((sc >> bit) & 1) {
      acc = AddXY<C>::run(acc, gx, gy);
      started = true;
    }
  }
  Fe<F> zi = fe_inv(acc.Z);
  uint4* o = tables + (size_t)t * (2 * F::N / 4);
  st_aos<F>(o, fe_mul(acc.X, zi));
  st_aos<F>(o + F::N / 4, fe_mul(acc.Y, zi));
}

constexpr int RP_BATCH = 8;

// Each thread produces RP_BATCH points (one shared inversion), canonical little-endian x | y.
template <class C>
__global__ void k_rp_points(const uint4* __restrict__ tables, uint8_t* __restrict__ out, size_t n, uint64_t seed,
                            size_t first) {
  using F = typename C::F;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t i0 = t * RP_BATCH;
  if (i0 >= n) return;
  typename C::Acc pts[RP_BATCH];
  Fe<F> pre[RP_BATCH];
  Fe<F> run = fe_one<F>();
  int m = (int)((n - i0 < (size_t)RP_BATCH) ? (n - i0) : RP_BATCH);
#pragma unroll 1
  for (int u = 0; u < m; u++) {
    uint64_t st = seed + (uint64_t)(first + i0 + u + 1) * 0xD1342543DE82EF95ull;  // `first`: global index of out[0]
    uint64_t r = splitmix64(st);
    typename C::Acc acc = C::zero();
#pragma unroll 1
    for (int k = 0; k < RP_TABLES; k++) {
      uint32_t w = (uint32_t)(r >> (RP_BITS * k)) & ((1u << RP_BITS) - 1u);
      const uint4* e = tables + ((size_t)k * (1u << RP_BITS) + w) * (2 * F::N / 4);
      acc = AddXY<C>::run(acc, ld_aos<F>(e), ld_aos<F>(e + F::N / 4));
    }
    if (fe_is_zero(acc.Z)) acc = AddXY<C>::run(C::zero(), fe_gen<F>(0), fe_gen<F>(1));
    pts[u] = acc;
    pre[u] = run;
    run = fe_mul(run, acc.Z);
  }
  Fe<F> inv = fe_inv(run);
  int nb = (F::BITS + 7) / 8;
#pragma unroll 1
  for (int u = m - 1; u >= 0; u--) {
    Fe<F> zi = fe_mul(inv, pre[u]);
    inv = fe_mul(inv, pts[u].Z);
    Fe<F> x = fe_from_mont(fe_mul(pts[u].X, zi));
    Fe<F> y = fe_from_mont(fe_mul(pts[u].Y, zi));
    uint8_t* o = out + (i0 + u) * (size_t)(2 * nb);
    for (int b = 0; b < nb; b++) {
      o[b] = (uint8_t)(x.v[b >> 2] >> (8 * (b & 3)));
      o[nb + b] = (uint8_t)(y.v[b >> 2] >> (8 * (b & 3)));
    }
  }
}

}  // namespace msm
