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
#include "inv_quad.cuh"

namespace msm {

constexpr int ACC_THREADS = 256;  // block size of the level-0 kernels
constexpr int ACC_MIN_PAIRS = 2;  // pairs per thread are chosen per round (RoundArgs::B0) between these
constexpr int ACC_MAX_PAIRS = 32; //   bounds so that the grid is about ACC_WAVES full waves of blocks
constexpr int ACC_WAVES = 8;      //   (dynamic block scheduling balances the SMs; few waves leave a tail)
constexpr int ACC_SINGLE_WAVE_MAX = 4 << 20;  // rounds with at most this many pairs run as one wave
constexpr int ACC_RESIDENT = 2;   // resident blocks per SM of k_bwd (126 registers x 256 threads), 12-limb fields
#ifndef MSM_ACC_RESIDENT_SMALL
#define MSM_ACC_RESIDENT_SMALL 3
#endif
// 8-limb fields: 3 blocks (<= 85 registers); their rounds lean on HBM latency, not only on the IMAD pipe
template <class F>
constexpr int acc_resident() {
  return F::N <= 8 ? MSM_ACC_RESIDENT_SMALL : ACC_RESIDENT;
}
constexpr int UP_THREADS = 64;    // block size of the serial product-tree levels
constexpr int UP_B1 = 8;          // elements per thread, serial levels
constexpr int TREE_CTA = 256;     // elements per block of the scan-based tree levels (one block per SM:
                                  // a modmul step costs 0.62 us x warps per sub-partition)
constexpr int TOP_CTA_MAX = 512;   // the top block handles up to this many elements
constexpr int TREE_MAX = 65536;   // serial levels until at most this many elements remain
constexpr int MAX_ROUNDS = 30;
constexpr int N_TOTALS = 2 * MAX_ROUNDS + 8;  // see k_scan
constexpr int FINISH_MAX_ELEMS = 16;  // the tree tail (k_finish) runs only with at most this many elements per bucket

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
  uint32_t* cnt;     // counts, indexed k * cnt_stride + bucket.  Classic: cnt_stride = L.  Shared buckets with few
                     //   buckets (L <= 2^16): still one counter per (window, bucket), merged into the L shared counts
                     //   afterwards (k_merge_counts) -- K times fewer atomics per address; with many buckets the
                     //   windows count straight into the shared array (cnt_stride = 0: K times fewer addresses)
  uint32_t* cursor;  // running positions (scatter), same indexing; start at 0, or (merged counts) at the window's
                     //   offset inside the shared bucket
  uint32_t cnt_stride;
  const uint32_t* po0;  // pair offsets of round 0
  uint32_t* ent;        // sorted entries, 2 * P0 slots
  uint32_t* pairkey;    // bucket of every round-0 pair
  uint32_t* digits;     // optional dump (tests): S*K
  // window k adds into buckets [k * bucket_stride, ...) and its entries point at base record
  // k * ent_stride + h.  Classic layout: (L, 0).  Shared buckets (precomputed 2^(kc) G tables, one
  // record set per window): (0, records per table) -- every window adds into the same L buckets.
  uint32_t bucket_stride;
  uint32_t ent_stride;
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
    const uint32_t kb = (uint32_t)k * a.cnt_stride + (l - 1);
    if (!SCATTER) {
      atomicAdd(&a.cnt[kb], 1u);
    } else {
      const uint32_t b = (uint32_t)k * a.bucket_stride + (l - 1);
      uint32_t pos = atomicAdd(&a.cursor[kb], 1u);
      uint32_t slot = 2u * a.po0[b] + pos;
      a.ent[slot] = ((uint32_t)h + (uint32_t)k * a.ent_stride) | ((carry ^ sign) << 31);  // (pairkey: k_fill_pairkey)
    }
  }
}

// Shared buckets: cnt[b] = sum over the windows of cntk[k * L + b], and cursor[k * L + b] = where window k's
// entries start inside bucket b.
static __global__ void __launch_bounds__(256) k_merge_counts(const uint32_t* __restrict__ cntk, int K, uint32_t L,
                                                            uint32_t* __restrict__ cnt, uint32_t* __restrict__ cursor) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= L) return;
  uint32_t run = 0;
  for (int k = 0; k < K; k++) {
    const uint32_t c = cntk[(size_t)k * L + b];
    cursor[(size_t)k * L + b] = run;
    run += c;
  }
  cnt[b] = run;
}

// Exclusive scan over buckets of the pair slots, for every tree round r at once (grid.y = round):
// pairs_r[b] = ceil(n_r[b] / 2) with n_r[b] = ceil(cnt[b] / 2^r) elements left -- but 0 once a bucket
// is down to one element (r >= 1): its sum then already sits in the `fin` array and it leaves the
// tree.  Two launches: per-tile sums (SCAN_TILE buckets per block), then every block adds the sums of
// the tiles before it and rescans its own tile.
//   po[r * NB + b], totals[r] = pair slots of round r, totals[MAX_ROUNDS+1] = max count,
//   totals[MAX_ROUNDS+2] = sum of counts, totals[MAX_ROUNDS+3+r] = additions performed in round r.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_TILE = SCAN_THREADS * 4;
constexpr int SCAN_ROUNDS_FIRST = 10;  // rounds scanned before the host knows the largest bucket (sizes up to 512);
                                       // the rest only when needed

// (`split` > 0, used by the generic bucket method: "round" 1 instead counts the virtual buckets
//  ceil(cnt / split) that an oversized bucket is cut into.)
__device__ __forceinline__ uint32_t scan_pairs_of(uint32_t c0, int r, uint32_t split = 0) {
  if (split && r == 1) return (c0 + split - 1) / split;
  uint32_t n = (uint32_t)(((unsigned long long)c0 + (1ull << r) - 1) >> r);
  return (n >= (r == 0 ? 1u : 2u)) ? ((n + 1) >> 1) : 0u;
}

// block-wide exclusive scan of one value per thread; returns the block total in `total`
__device__ __forceinline__ uint32_t scan_block_exclusive(uint32_t local, uint32_t& total, uint32_t* wsum) {
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  uint32_t incl = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t x = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0u, xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, xi, d);
      if (lane >= d) xi += o;
    }
    wsum[lane] = xi - x;
    if (lane == 31) wsum[32] = xi;
  }
  __syncthreads();
  total = wsum[32];
  return wsum[w] + incl - local;
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t* __restrict__ cnt, uint32_t NB,
                                                                   uint32_t* __restrict__ tilesum, uint32_t ntiles,
                                                                   unsigned long long* __restrict__ totals, uint32_t split,
                                                                   int r0) {
  __shared__ uint32_t wsum[33];
  const int r = r0 + blockIdx.y;
  const uint32_t b0 = blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
  uint32_t local = 0, mx = 0;
  unsigned long long nent = 0, nadd = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t c0 = (b0 + j < NB) ? cnt[b0 + j] : 0u;
    mx = max(mx, c0);
    nent += c0;
    nadd += (uint32_t)(((unsigned long long)c0 + (1ull << r) - 1) >> r) >> 1;
    local += scan_pairs_of(c0, r, split);
  }
  uint32_t total;
  scan_block_exclusive(local, total, wsum);
  if (threadIdx.x == 0) tilesum[(size_t)r * ntiles + blockIdx.x] = total;
  // block-reduce the statistics through the same shared array
  for (int d = 16; d >= 1; d >>= 1) {
    mx = max(mx, __shfl_down_sync(0xffffffffu, mx, d));
    nent += __shfl_down_sync(0xffffffffu, nent, d);
    nadd += __shfl_down_sync(0xffffffffu, nadd, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&totals[MAX_ROUNDS + 3 + r], nadd);
    if (r == 0) {
      atomicMax(&totals[MAX_ROUNDS + 1], (unsigned long long)mx);
      atomicAdd(&totals[MAX_ROUNDS + 2], nent);
    }
  }
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_write(const uint32_t* __restrict__ cnt, uint32_t NB,
                                                                   const uint32_t* __restrict__ tilesum, uint32_t ntiles,
                                                                   uint32_t* __restrict__ po,
                                                                   unsigned long long* __restrict__ totals, uint32_t split,
                                                                   int r0) {
  __shared__ uint32_t wsum[33];
  __shared__ unsigned long long sbase;
  const int r = r0 + blockIdx.y;
  // base = sum of the tile sums before this tile
  unsigned long long part = 0;
  for (uint32_t i = threadIdx.x; i < blockIdx.x; i += SCAN_THREADS) part += tilesum[(size_t)r * ntiles + i];
  for (int d = 16; d >= 1; d >>= 1) part += __shfl_down_sync(0xffffffffu, part, d);
  if (threadIdx.x == 0) sbase = 0;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) atomicAdd(&sbase, part);
  __syncthreads();
  const unsigned long long base = sbase;
  const uint32_t b0 = blockIdx.x * SCAN_TILE + 4 * threadIdx.x;
  uint32_t v[4], local = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    v[j] = scan_pairs_of((b0 + j < NB) ? cnt[b0 + j] : 0u, r, split);
    local += v[j];
  }
  uint32_t total;
  unsigned long long off = base + scan_block_exclusive(local, total, wsum);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (b0 + j < NB) po[(size_t)r * NB + b0 + j] = (uint32_t)off;
    off += v[j];
  }
  if (blockIdx.x == ntiles - 1 && threadIdx.x == 0) totals[r] = base + total;
}

// pairkey[i] = bucket of round-0 pair i, written as coalesced runs: one warp per 32 consecutive
// buckets, lanes stride over each bucket's pair range (replaces a second scattered 4-byte write per
// entry in the scatter pass).
static __global__ void __launch_bounds__(256) k_fill_pairkey(const uint32_t* __restrict__ po0, uint32_t NB,
                                                            const unsigned long long* __restrict__ totals,
                                                            uint32_t* __restrict__ pairkey) {
  const uint32_t P0 = (uint32_t)totals[0];  // pair slots of round 0 (the host may not know it yet)
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b0 = warp * 32;
  if (b0 >= NB) return;
  const uint32_t mine = (b0 + lane < NB) ? po0[b0 + lane] : P0;  // start of bucket b0 + lane
  const uint32_t end = (b0 + 32 < NB) ? po0[b0 + 32] : P0;
#pragma unroll 1
  for (int src = 0; src < 32; src++) {
    const uint32_t st = __shfl_sync(0xffffffffu, mine, src);
    const uint32_t en = (src < 31) ? __shfl_sync(0xffffffffu, mine, (src + 1) & 31) : end;
    for (uint32_t i = st + lane; i < en; i += 32) pairkey[i] = b0 + src;
  }
}

// ------------------------------------------------------------------------------------------
// one round of the pairwise tree
// ------------------------------------------------------------------------------------------
// final bucket sums (affine, infinity-marked when the bucket cancelled out), indexed by bucket
template <class F>
struct FinBuf {
  uint4* base;
  size_t cap;
  static constexpr int CH = F::N / 4;
  __device__ __forceinline__ void store(size_t b, const Aff<F>& P) const {
    st_soa<F>(base, cap, b, P.x);
    st_soa<F>(base + (size_t)CH * cap, cap, b, P.y);
  }
  __device__ __forceinline__ Aff<F> load(size_t b) const {
    Aff<F> P;
    P.x = ld_soa<F>(base, cap, b);
    P.y = ld_soa<F>(base + (size_t)CH * cap, cap, b);
    return P;
  }
  static size_t bytes(size_t cap) { return (size_t)2 * CH * cap * sizeof(uint4); }
};

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
  FinBuf<F> fin;         // buckets that are down to one element
  // batch inversion, level 0
  uint4* prefix;         // P elements, stride = P
  uint4* tot;            // one per thread, stride = M1: the thread's total (BLK: product of the OTHER
                         //   threads' totals of its block)
  const uint4* invtot;   // inverse of tot, same layout (BLK: inverse of the block totals, one per block)
  uint4* blktot;         // BLK: one per block, stride = gridDim.x
  size_t M1;
  int B0;                // pairs per thread: a block owns ACC_THREADS * B0 consecutive pairs
};

// ---- scan-based top of the product tree (latency: ~25 dependent modmuls for a 1024x reduction
// instead of 3 per element in a serial chain) ------------------------------------------------
template <class F>
__device__ __forceinline__ Fe<F> fe_shfl_up(const Fe<F>& v, int d) {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = __shfl_up_sync(0xffffffffu, v.v[i], d);
  return r;
}
template <class F>
__device__ __forceinline__ Fe<F> fe_shfl_down(const Fe<F>& v, int d) {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = __shfl_down_sync(0xffffffffu, v.v[i], d);
  return r;
}
template <class F>
__device__ __forceinline__ Fe<F> fe_shfl(const Fe<F>& v, int src) {
  Fe<F> r;
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src);
  return r;
}
template <class F>
__device__ __forceinline__ void fe_to_smem(uint32_t* s, const Fe<F>& v) {
#pragma unroll
  for (int i = 0; i < F::N; i++) s[i] = v.v[i];
}
template <class F>
__device__ __forceinline__ Fe<F> fe_from_smem(const uint32_t* s) {
  Fe<F> v;
#pragma unroll
  for (int i = 0; i < F::N; i++) v.v[i] = s[i];
  return v;
}

// inclusive prefix (p) and suffix (s) products across the lanes of a warp
template <class F>
__device__ __forceinline__ void warp_scan_products(const Fe<F>& v, int lane, Fe<F>& p, Fe<F>& s) {
  p = v;
  s = v;
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    Fe<F> tp = fe_shfl_up(p, d), ts = fe_shfl_down(s, d);
    Fe<F> np = fe_mul(p, tp), ns = fe_mul(s, ts);
    p = fe_select(lane >= d, np, p);
    s = fe_select(lane + d < 32, ns, s);
  }
}

// Products over a block of NT threads (one value per thread): `others` = product of all the
// block's values but this thread's, `total` = product of all of them (valid in every thread).
// ~13 dependent modmuls per thread (two 5-step warp scans + 3) and one more scan in warp 0.
template <class F, int NT>
__device__ __forceinline__ void block_products(const Fe<F>& v, Fe<F>& others, Fe<F>& total, uint32_t* smem) {
  constexpr int NW = NT / 32;
  uint32_t* wtot = smem;
  uint32_t* wpre = smem + 32 * F::N;
  uint32_t* wsuf = smem + 64 * F::N;
  uint32_t* btot = smem + 96 * F::N;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  Fe<F> p, s;
  warp_scan_products(v, lane, p, s);
  if (lane == 31) fe_to_smem<F>(wtot + w * F::N, p);
  __syncthreads();
  if (w == 0) {
    Fe<F> x = (lane < NW) ? fe_from_smem<F>(wtot + lane * F::N) : fe_one<F>(), xp, xs;
    warp_scan_products(x, lane, xp, xs);
    Fe<F> pre = fe_shfl_up(xp, 1), suf = fe_shfl_down(xs, 1);
    if (lane == 0) pre = fe_one<F>();
    if (lane == 31) suf = fe_one<F>();
    fe_to_smem<F>(wpre + lane * F::N, pre);
    fe_to_smem<F>(wsuf + lane * F::N, suf);
    if (lane == 31) fe_to_smem<F>(btot, xp);
  }
  __syncthreads();
  Fe<F> pe = fe_shfl_up(p, 1), se = fe_shfl_down(s, 1);
  if (lane == 0) pe = fe_one<F>();
  if (lane == 31) se = fe_one<F>();
  others = fe_mul(fe_mul(fe_from_smem<F>(wpre + w * F::N), pe), fe_mul(se, fe_from_smem<F>(wsuf + w * F::N)));
  total = fe_from_smem<F>(btot);
}
constexpr int BLOCK_PRODUCTS_SMEM_WORDS(int n) { return 97 * n; }

// Same result with two thirds of the field products, for kernels in which EVERY resident warp runs this at the
// same time (k_fwd in block mode: the scans are then bound by the IMAD pipe, not by latency): neighbouring
// lanes multiply their values first, only NT / 2 threads (half of the warps) scan the pair products, and each
// thread finishes with one product by its neighbour's value.  `stage`: NT / 2 field elements of shared memory.
template <class F, int NT>
__device__ __forceinline__ void block_products_paired(const Fe<F>& v, Fe<F>& others, Fe<F>& total, uint32_t* smem,
                                                      uint32_t* stage) {
  constexpr int NW = NT / 64;  // warps that scan
  uint32_t* wtot = smem;
  uint32_t* wpre = smem + 32 * F::N;
  uint32_t* wsuf = smem + 64 * F::N;
  uint32_t* btot = smem + 96 * F::N;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const Fe<F> partner = fe_shfl(v, lane ^ 1);
  const Fe<F> pp = fe_mul(v, partner);
  if (!(t & 1)) fe_to_smem<F>(stage + (t >> 1) * F::N, pp);
  __syncthreads();
  const bool active = t < NT / 2;  // whole warps
  Fe<F> p, s;
  if (active) {
    warp_scan_products(fe_from_smem<F>(stage + t * F::N), lane, p, s);
    if (lane == 31) fe_to_smem<F>(wtot + w * F::N, p);
  }
  __syncthreads();
  if (w == 0) {
    Fe<F> x = (lane < NW) ? fe_from_smem<F>(wtot + lane * F::N) : fe_one<F>(), xp, xs;
    warp_scan_products(x, lane, xp, xs);
    Fe<F> pre = fe_shfl_up(xp, 1), suf = fe_shfl_down(xs, 1);
    if (lane == 0) pre = fe_one<F>();
    if (lane == 31) suf = fe_one<F>();
    fe_to_smem<F>(wpre + lane * F::N, pre);
    fe_to_smem<F>(wsuf + lane * F::N, suf);
    if (lane == 31) fe_to_smem<F>(btot, xp);
  }
  __syncthreads();
  if (active) {
    Fe<F> pe = fe_shfl_up(p, 1), se = fe_shfl_down(s, 1);
    if (lane == 0) pe = fe_one<F>();
    if (lane == 31) se = fe_one<F>();
    Fe<F> o = fe_mul(fe_mul(fe_from_smem<F>(wpre + w * F::N), pe), fe_mul(se, fe_from_smem<F>(wsuf + w * F::N)));
    fe_to_smem<F>(stage + t * F::N, o);  // product of all OTHER pairs, in place of this thread's own pair product
  }
  __syncthreads();
  others = fe_mul(fe_from_smem<F>(stage + (t >> 1) * F::N), partner);
  total = fe_from_smem<F>(btot);
}

// Loads of one pair.  Everything that depends only on the pair index (the two elements, or the two
// sorted entries and then the gathered base points) is issued BEFORE the bucket lookups
// (pairkey -> po_r / cnt), so a pair costs two dependent memory round trips instead of four.  The
// second slot of a bucket's last pair may be padding: it is read anyway (the slot exists; round-0
// entry padding is zeroed) and discarded.
template <class F, bool R0>
__device__ __forceinline__ void load_pair(const RoundArgs<F>& a, size_t i, Aff<F>& A, Aff<F>& B, uint32_t& b,
                                          uint32_t& j) {
  if (R0) {
    uint2 e = reinterpret_cast<const uint2*>(a.ent)[i];
    b = a.pairkey[i];
    A = gather_base<F>(a.bases, e.x);
    B = gather_base<F>(a.bases, e.y);
  } else {
    b = a.pairkey[i];
    A = a.in.load(2 * i);
    B = a.in.load(2 * i + 1);
  }
  j = (uint32_t)i - a.po_r[b];
  uint32_t n = (uint32_t)(((unsigned long long)a.cnt[b] + (1ull << a.r) - 1) >> a.r);
  if (!(2 * j + 1 < n)) B = aff_inf<F>();
}

// Denominator of pair i for the forward pass.  Only the x coordinates are needed unless they are
// equal (doubling / inverse pair), so the common path reads 2 x 48 bytes instead of 2 x 96: the
// forward pass is HBM-bound (1 modmul per ~150 bytes), the backward pass is not.
template <class F, bool R0>
__device__ __forceinline__ bool fwd_denominator(const RoundArgs<F>& a, size_t i, Fe<F>& d) {
  Fe<F> xa, xb;
  uint2 e = make_uint2(0u, 0u);
  uint32_t b;
  if (R0) {
    e = reinterpret_cast<const uint2*>(a.ent)[i];
    b = a.pairkey[i];
    xa = ld_aos<F>(a.bases + (size_t)ent_index(e.x) * (2 * F::N / 4));
    xb = ld_aos<F>(a.bases + (size_t)ent_index(e.y) * (2 * F::N / 4));
  } else {
    b = a.pairkey[i];
    xa = ld_soa<F>(a.in.coord(0, 0), a.in.cap, i);
    xb = ld_soa<F>(a.in.coord(1, 0), a.in.cap, i);
  }
  uint32_t j = (uint32_t)i - a.po_r[b];
  uint32_t n = (uint32_t)(((unsigned long long)a.cnt[b] + (1ull << a.r) - 1) >> a.r);
  if (!(2 * j + 1 < n)) return false;  // single element: passes through
  if (xa.v[F::N - 1] == AFF_INF_MARK || xb.v[F::N - 1] == AFF_INF_MARK) return false;
  d = fe_sub(xb, xa);
  if (!fe_is_zero(d)) return true;
  Aff<F> A, B;
  if (R0) {
    A = gather_base<F>(a.bases, e.x);
    B = gather_base<F>(a.bases, e.y);
  } else {
    A = a.in.load(2 * i);
    B = a.in.load(2 * i + 1);
  }
  return aff_add_prepare(A, B, d) <= AFF_DBL;
}

// forward pass: exclusive prefix products of the denominators, per thread.
// A block owns ACC_THREADS * B0 consecutive pairs; thread t takes pairs t, t + 256, ... of them, so
// every warp access is coalesced.
// BLK (small rounds): the block multiplies its thread totals on the spot (block_products), so the
// product tree starts from one element per block and two kernel launches per round disappear.
template <class F, bool R0, bool BLK>
__global__ void __launch_bounds__(ACC_THREADS, acc_resident<F>()) k_fwd(RoundArgs<F> a) {
  __shared__ uint32_t smem[BLK ? 97 * F::N : 1];
  __shared__ uint32_t stage[BLK ? (ACC_THREADS / 2) * F::N : 1];
  const size_t chunk0 = (size_t)blockIdx.x * ((size_t)ACC_THREADS * a.B0);
  const size_t gid = (size_t)blockIdx.x * ACC_THREADS + threadIdx.x;
  Fe<F> run = fe_one<F>();
#pragma unroll 1
  for (int s = 0; s < a.B0; s++) {
    const size_t i = chunk0 + (size_t)s * ACC_THREADS + threadIdx.x;
    if (i >= a.P) break;
    Fe<F> d;
    if (fwd_denominator<F, R0>(a, i, d)) {
      st_soa<F>(a.prefix, a.P, i, run);
      run = fe_mul(run, d);
    }
  }
  if (BLK) {
    Fe<F> others, total;
    block_products_paired<F, ACC_THREADS>(run, others, total, smem, stage);
    st_soa<F>(a.tot, a.M1, gid, others);
    if (threadIdx.x == 0) st_soa<F>(a.blktot, gridDim.x, blockIdx.x, total);
  } else {
    st_soa<F>(a.tot, a.M1, gid, run);
  }
}

// backward pass: individual inverses from the running inverse, then finish the additions
template <class F, bool R0, bool BLK>
__global__ void __launch_bounds__(ACC_THREADS, acc_resident<F>()) k_bwd(RoundArgs<F> a) {
  const size_t chunk0 = (size_t)blockIdx.x * ((size_t)ACC_THREADS * a.B0);
  const size_t gid = (size_t)blockIdx.x * ACC_THREADS + threadIdx.x;
  Fe<F> inv = BLK ? fe_mul(ld_soa<F>(a.invtot, gridDim.x, blockIdx.x), ld_soa<F>(a.tot, a.M1, gid))
                  : ld_soa<F>(a.invtot, a.M1, gid);
  // same pairs as the forward pass, in reverse order
#pragma unroll 1
  for (int s = a.B0 - 1; s >= 0; s--) {
    const size_t i = chunk0 + (size_t)s * ACC_THREADS + threadIdx.x;
    if (i >= a.P) continue;
    Aff<F> A, B;
    uint32_t b, j;
    Fe<F> pre = ld_soa<F>(a.prefix, a.P, i);  // (unused garbage for pass-through pairs)
    load_pair<F, R0>(a, i, A, B, b, j);
    // round 0 (gathered operands: the tightest register budget) looks the output slot up after the arithmetic
    // instead of keeping it live across it (it was the 20 bytes that spilled at 128 registers)
    uint32_t pon = R0 ? 0u : a.po_n[b];
    Fe<F> d;
    int cs = aff_add_prepare(A, B, d);
    Fe<F> id = inv;
    if (cs <= AFF_DBL) {
      id = fe_mul(inv, pre);
      inv = fe_mul(inv, d);
    }
    Aff<F> R = aff_add_finish(cs, A, B, id);
    uint32_t n = (uint32_t)(((unsigned long long)a.cnt[b] + (1ull << a.r) - 1) >> a.r);
    if (n <= 2) {  // one element left after this round: the bucket sum
      a.fin.store(b, R);
    } else {
      if (R0) {
        pon = a.po_n[b];
        j = (uint32_t)i - a.po_r[b];
      }
      size_t e = 2 * (size_t)pon + j;
      a.out.store(e, R);
      if (!(e & 1)) a.pairkey_next[e >> 1] = b;
    }
  }
}

// Tail of the tree.  Once the additions that are left cost less as projective mixed additions than
// as further latency-bound rounds (each round pays a product tree + one inversion), every unfinished
// bucket is summed by one thread and ALL bucket sums are written out projective (no inversion); the
// bucket reduction then reads this array instead of `fin`.  Two kernels:
//   k_finish_slots  one thread per PAIR SLOT of the round that would come next; the thread that owns a
//                   bucket's first slot sums the bucket's elements.  Unfinished buckets are contiguous in
//                   slot space, so the warps doing the mixed additions are densely populated (one thread
//                   per bucket over all NB buckets leaves half the lanes of every warp idle);
//   k_finish_rest   one thread per bucket for the cheap cases: empty -> neutral element, finished in an
//                   earlier round -> copied from `fin`.
template <class F, uint32_t B3, bool R0>
__global__ void __launch_bounds__(64) k_finish_slots(RoundArgs<F> a, uint4* __restrict__ buckets) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.P) return;
  const uint32_t b = a.pairkey[i];
  if ((uint32_t)i != a.po_r[b]) return;  // not the bucket's first slot
  const uint32_t n = (uint32_t)(((unsigned long long)a.cnt[b] + (1ull << a.r) - 1) >> a.r);
  const size_t e0 = 2 * i;
  Proj<F> acc = proj_from_aff(R0 ? gather_base<F>(a.bases, a.ent[e0]) : a.in.load(e0));
#pragma unroll 1
  for (uint32_t j = 1; j < n; j++) {
    Aff<F> Q = R0 ? gather_base<F>(a.bases, a.ent[e0 + j]) : a.in.load(e0 + j);
    if (aff_is_inf(Q)) continue;
    if constexpr (F::LAZY && B3 == 3) acc = proj_add_mixed_nr<F>(acc, Q);  // coordinates < 2p in registers
    else acc = proj_add_mixed<F, B3>(acc, Q);
  }
  if constexpr (F::LAZY && B3 == 3) acc = proj_canon(acc);
  uint4* o = buckets + (size_t)b * (3 * F::N / 4);
  st_aos<F>(o, acc.X);
  st_aos<F>(o + F::N / 4, acc.Y);
  st_aos<F>(o + 2 * F::N / 4, acc.Z);
}

template <class F, bool R0>
__global__ void __launch_bounds__(128) k_finish_rest(RoundArgs<F> a, uint32_t NB, uint4* __restrict__ buckets) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= NB) return;
  const uint32_t cnt = a.cnt[b];
  const uint32_t n = (uint32_t)(((unsigned long long)cnt + (1ull << a.r) - 1) >> a.r);
  Proj<F> acc;
  if (cnt == 0) acc = proj_zero<F>();
  else if (!R0 && n == 1) acc = proj_from_aff(a.fin.load(b));  // finished in an earlier round
  else return;                                                 // has pair slots: k_finish_slots
  uint4* o = buckets + (size_t)b * (3 * F::N / 4);
  st_aos<F>(o, acc.X);
  st_aos<F>(o + F::N / 4, acc.Y);
  st_aos<F>(o + 2 * F::N / 4, acc.Z);
}

// Same tail for dense buckets (shared-bucket mode: every bucket still has ~2^k elements when the tree stops,
// so slot-indexed threads would leave most lanes of a warp idle): TWO adjacent lanes per bucket, each sums
// half of the bucket's elements with mixed additions, then one complete addition joins the halves (the serial
// chain is half as long and twice as many warps hide its latency: 210 -> ~120 us for 2^15 buckets of ~8).
template <class F, uint32_t B3, bool R0>
__global__ void __launch_bounds__(64) k_finish_buckets(RoundArgs<F> a, uint32_t NB, uint4* __restrict__ buckets) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b = t >> 1, half = t & 1u;
  const bool live = b < NB;  // both lanes of a pair are live or dead together
  const uint32_t cnt = live ? a.cnt[b] : 0u;
  const uint32_t n = (uint32_t)(((unsigned long long)cnt + (1ull << a.r) - 1) >> a.r);
  Proj<F> acc = proj_zero<F>();
  if (cnt != 0 && !R0 && n == 1) {
    if (half == 0) acc = proj_from_aff(a.fin.load(b));  // finished in an earlier round
  } else if (cnt != 0) {
    const size_t e0 = 2 * (size_t)a.po_r[b];
    const uint32_t mid = (n + 1) >> 1;
    const uint32_t lo = half ? mid : 0u, hi = half ? n : mid;
    if (lo < hi) acc = proj_from_aff(R0 ? gather_base<F>(a.bases, a.ent[e0 + lo]) : a.in.load(e0 + lo));
#pragma unroll 1
    for (uint32_t j = lo + 1; j < hi; j++) {
      Aff<F> Q = R0 ? gather_base<F>(a.bases, a.ent[e0 + j]) : a.in.load(e0 + j);
      if (aff_is_inf(Q)) continue;
      if constexpr (F::LAZY && B3 == 3) acc = proj_add_mixed_nr<F>(acc, Q);
      else acc = proj_add_mixed<F, B3>(acc, Q);
    }
  }
  // join the halves: complete addition with the partner lane's sum (neutral element where a half is empty)
  Proj<F> other;
#pragma unroll
  for (int i = 0; i < F::N; i++) {
    other.X.v[i] = __shfl_xor_sync(0xffffffffu, acc.X.v[i], 1);
    other.Y.v[i] = __shfl_xor_sync(0xffffffffu, acc.Y.v[i], 1);
    other.Z.v[i] = __shfl_xor_sync(0xffffffffu, acc.Z.v[i], 1);
  }
  if constexpr (F::LAZY && B3 == 3) acc = proj_canon(proj_add_nr<F>(acc, other));
  else acc = proj_add<F, B3>(acc, other);
  if (!live || half) return;
  uint4* o = buckets + (size_t)b * (3 * F::N / 4);
  st_aos<F>(o, acc.X);
  st_aos<F>(o + F::N / 4, acc.Y);
  st_aos<F>(o + 2 * F::N / 4, acc.Z);
}

// Window tables for RESIDENT bases (built once per point set by set_bases): table k holds 2^(kc) G_i and
// its endomorphism image for every base point, so that the digits of ALL windows can be added into ONE
// set of 2^(c-1) buckets -- the bucket reduction then runs over L instead of K L buckets and the Horner
// combination of the windows (K - 1) c dependent doublings in one warp) disappears.  One thread per
// point: c projective doublings of the previous table's point, one inversion per block (product of the
// block's Z values, block_products), affine records out.  The reference has no counterpart (its points are
// copied and sign-folded on every call, src/msm-batched-affine.ts:338-409); the sum is the same group
// element.
template <class F, uint32_t B3>
__global__ void __launch_bounds__(256) k_build_table(const uint4* __restrict__ prev, uint4* __restrict__ next, size_t n, int c) {
  __shared__ uint32_t smem[97 * F::N + F::N];
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  const bool valid = i < n;
  Aff<F> P = aff_inf<F>();
  if (valid) {
    const uint4* r = prev + i * (size_t)(4 * F::N / 4);
    P.x = ld_aos<F>(r);
    P.y = ld_aos<F>(r + F::N / 4);
  }
  Proj<F> A = proj_from_aff(P);
#pragma unroll 1
  for (int d = 0; d < c; d++) A = proj_dbl<F, B3>(A);
  const bool inf = fe_is_zero(A.Z);
  Fe<F> others, total;
  block_products<F, 256>(inf ? fe_one<F>() : A.Z, others, total, smem);
  uint32_t* binv = smem + 97 * F::N;
  __syncthreads();
  if (threadIdx.x == 0) fe_to_smem<F>(binv, fe_inv(total));
  __syncthreads();
  if (!valid) return;
  Aff<F> Q = aff_inf<F>(), E = Q;
  if (!inf) {
    const Fe<F> zi = fe_mul(others, fe_from_smem<F>(binv));
    Q.x = fe_mul(A.X, zi);
    Q.y = fe_mul(A.Y, zi);
    E.x = fe_mul(Q.x, fe_beta<F>());
    E.y = Q.y;
  }
  uint4* o = next + i * (size_t)(4 * F::N / 4);
  st_aos<F>(o, Q.x);
  st_aos<F>(o + F::N / 4, Q.y);
  st_aos<F>(o + 2 * F::N / 4, E.x);
  st_aos<F>(o + 3 * F::N / 4, E.y);
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

// One block handles CTA elements: `others[i]` = product of all the block's elements but i,
// `tot[block]` = product of all of them.  TOP: the block is the whole level; it inverts its total
// and writes the individual inverses straight to `others`.
template <class F, bool TOP, int CTA>
__global__ void __launch_bounds__(CTA) k_tree_up(const uint4* __restrict__ val, size_t M, uint4* __restrict__ others,
                                                 uint4* __restrict__ tot, size_t Mn) {
  __shared__ uint32_t smem[97 * F::N + F::N];
  const size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
  Fe<F> v = (i < M) ? ld_soa<F>(val, M, i) : fe_one<F>();
  Fe<F> o, total;
  block_products<F, CTA>(v, o, total, smem);
  if (TOP) {
    uint32_t* binv = smem + 97 * F::N;
    __syncthreads();
    if (threadIdx.x == 0) fe_to_smem<F>(binv, fe_inv(total));
    __syncthreads();
    o = fe_mul(o, fe_from_smem<F>(binv));
  } else if (threadIdx.x == 0) {
    st_soa<F>(tot, Mn, blockIdx.x, total);
  }
  if (i < M) st_soa<F>(others, M, i, o);
}

// Top of the product tree for up to 2 * CTA values: ONE block, two values per thread, so that the
// shuffle scans run with half as many warps per SM sub-partition (the scans are bound by the IMAD pipe of
// the one SM they run on).  inverses[i] = 1 / val[i].
template <class F, int CTA>
__global__ void __launch_bounds__(CTA) k_tree_top2(const uint4* __restrict__ val, size_t M, uint4* __restrict__ inverses) {
  __shared__ uint32_t smem[97 * F::N + F::N];
  const size_t i0 = (size_t)threadIdx.x * 2;
  Fe<F> v0 = (i0 < M) ? ld_soa<F>(val, M, i0) : fe_one<F>();
  Fe<F> v1 = (i0 + 1 < M) ? ld_soa<F>(val, M, i0 + 1) : fe_one<F>();
  Fe<F> o, total;
  block_products<F, CTA>(fe_mul(v0, v1), o, total, smem);
  uint32_t* binv = smem + 97 * F::N;
  __syncthreads();
  if (threadIdx.x < 32) {  // warp 0 inverts the grand total together (inv_quad.cuh)
    const Fe<F> inv = fe_inv_quad(total);
    if (threadIdx.x == 0) fe_to_smem<F>(binv, inv);
  }
  __syncthreads();
  o = fe_mul(o, fe_from_smem<F>(binv));  // 1 / (v0 v1)
  if (i0 < M) st_soa<F>(inverses, M, i0, fe_mul(o, v1));
  if (i0 + 1 < M) st_soa<F>(inverses, M, i0 + 1, fe_mul(o, v0));
}

// others[i] <- invtot[block] * others[i]  = 1 / val[i]
template <class F>
__global__ void __launch_bounds__(TREE_CTA) k_tree_down(uint4* __restrict__ others, size_t M, const uint4* __restrict__ invtot,
                                                        size_t Mn) {
  const size_t i = (size_t)blockIdx.x * TREE_CTA + threadIdx.x;
  if (i >= M) return;
  Fe<F> it = ld_soa<F>(invtot, Mn, blockIdx.x);
  st_soa<F>(others, M, i, fe_mul(it, ld_soa<F>(others, M, i)));
}

}  // namespace msm
