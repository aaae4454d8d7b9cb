// Host orchestration templates of libmsm_b200.so (instantiated once per curve in curve_*.cu).
// One context = one GPU + one curve.  All device work goes to ctx->stream.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/msm_b200.h"
#include "kernels_weierstrass.cuh"
#include "kernels_basic.cuh"

using namespace msm;

std::string& msm_global_err();
#define g_err (msm_global_err())

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

// Phase marks of the last MSM, resolved into a msm_b200_timing after the stream has drained: at once when the
// caller passed a timing struct, or later through msm_b200_last_timing() so that the call itself need not
// synchronise (the multi-GPU driver queues its collective behind the MSM first).
struct PendingTiming {
  bool valid = false;
  int e[5] = {0, 0, 0, 0, 0};  // digits | sort | accumulate | reduce boundaries
  std::vector<std::pair<int, int>> hot;
  int window_bits = 0, n_windows = 0, rounds = 0, shared_buckets = 0;
  unsigned long long n_adds = 0;
  int h2d[2] = {-1, -1};  // scalar upload, set by the entry point
  int fwd0[2] = {-1, -1}; // round-0 forward pass (k_fwd over the gathered base points)
  unsigned long long fwd0_pairs = 0;
};

struct msm_b200_ctx {
  int device = 0;
  int curve = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // one-shot call: points are copied and ingested on a second stream while the scalars are already
  // being decomposed and sorted; the first kernel that reads the bases waits for this event
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t bases_ready = nullptr;
  cudaEvent_t totals_ready = nullptr;  // bucket totals have reached the pinned host buffer
  // host scalars arrive in SC_CHUNKS pieces on the copy stream; the per-scalar phases (GLV / unpacking, digit
  // histogram) run piece by piece behind them, so all but the last piece of the upload is hidden
  static constexpr int SC_CHUNKS = 4;
  cudaEvent_t sc_ready[SC_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t sc_start = nullptr;
  int sc_chunks = 0;  // > 0: the current run's scalars come in that many pieces (consumed by the run)
  bool bases_pending = false;
  std::string err;
  int launches = 0;
  int sm_count = 148;
  // cost model of the tree tail (run_affine_glv); MSM_B200_FINISH_ADD / MSM_B200_FINISH_ROUND override (tuning)
  double finish_add_modmuls = 18.0;
  double finish_round_modmuls = 0;  // 0: 1e6 (classic layout, slot-indexed tail) / 3e6 (shared buckets, dense tail)
  int finish_max_elems = 0;         // 0: 16 / 32 elements per bucket at most when the tail takes over
  int acc_min_pairs = ACC_MIN_PAIRS;  // MSM_B200_ACC_MIN_PAIRS (tuning)
  // bits per level of the bucket reduction; 0 = by bucket count (reduce_buckets): 2^18 buckets: 8 buckets per
  // thread at level 0, one item per lane while > 4096 items; <= 2^16 buckets (shared-bucket mode): level 0
  // with one lane quad per 4 buckets, every further level quad-cooperative (tools/sweep_reduce.sh)
  int reduce_gb0 = 0;          // MSM_B200_REDUCE_GB0 (tuning)
  int reduce_warp_gb = 3;      // MSM_B200_REDUCE_WARP_GB (2^18 buckets: 5 -> 1.29 ms, 4 -> 1.25, 3 -> 1.22, 2 -> 1.24)
  size_t reduce_warp_min = 0;  // MSM_B200_REDUCE_WARP_MIN: levels with more items use one lane per item
  int reduce_quad0 = -1;       // MSM_B200_REDUCE_Q0: level 0 on lane quads (1) or lone lanes (0); -1 = by bucket count
  // resident bases
  DevBuf bases;
  size_t n_bases = 0;
  // bases borrowed from another context of the same device and curve (msm_b200_share_bases): several contexts
  // can then run MSMs over ONE resident point set at the same time (pipelined MSMs, see bench.py `pipelined`).
  // `bases_gen` counts this context's set_bases calls; a borrower remembers the lender's value.
  msm_b200_ctx* bases_owner = nullptr;
  unsigned long long bases_gen = 0, borrowed_gen = 0;
  std::vector<msm_b200_ctx*> borrowers;  // contexts that read this context's bases (told when it is destroyed)
  // window tables of the resident bases (shared-bucket mode, see k_build_table): table k = 2^(k * table_c) G
  int table_c = 0, table_K = 0;      // 0: no tables
  bool tables_enabled = true;        // MSM_B200_TABLES=0 disables
  int table_max_log2n = 25;          // largest point set that gets tables (MSM_B200_TABLE_MAX_LOG2N): 2^26 points
                                     // with 6 record sets + the workspace of the rounds exceed 180 GB
  int table_window = 0;              // MSM_B200_TABLE_WINDOW: forces the tables' window size (tuning)
  int bucket_split = 0;              // MSM_B200_BUCKET_SPLIT: entries per virtual bucket of the generic bucket method (0 = by size)
  // workspace
  DevBuf raw_points, raw_scalars, hs, cnt, cntk, cursor, po, totals, ent, pairkey[2], elem[2], prefix;
  DevBuf lvl_pre[8], lvl_tot[8], red[2], partial, result, buckets, rp_tables, fin, others, tilesum;
  unsigned long long* h_totals = nullptr;  // pinned
  uint32_t* h_result = nullptr;            // pinned
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  PendingTiming pending;
};

#define CK(call)                                                                             \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      cudaGetLastError(); /* a non-sticky error must not surface again in a later, unrelated call */ \
      char buf_[512];                                                                        \
      snprintf(buf_, sizeof buf_, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      if (ctx) ctx->err = buf_;                                                              \
      g_err = buf_;                                                                          \
      return (e_ == cudaErrorMemoryAllocation) ? MSM_E_NOMEM : MSM_E_CUDA;                   \
    }                                                                                        \
  } while (0)

#define RET_IF(x)           \
  do {                      \
    int rc_ = (x);          \
    if (rc_ != 0) return rc_; \
  } while (0)

static int fail(msm_b200_ctx* ctx, int code, const char* msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

static int ensure(msm_b200_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (b.cap >= bytes) return 0;
  if (b.p) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
  }
  size_t want = bytes + bytes / 16;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    CK(cudaMalloc(&b.p, want));
  }
  b.cap = want;
  return 0;
}

static void release(DevBuf& b) {
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
}

#define LAUNCH(ctx, kern, grid, block, ...)                          \
  do {                                                               \
    kern<<<(grid), (block), 0, (ctx)->stream>>>(__VA_ARGS__);        \
    (ctx)->launches++;                                               \
  } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// the resident record sets this context reads (its own or the lender's)
static inline const uint4* bases_ptr(const msm_b200_ctx* ctx) {
  return (const uint4*)(ctx->bases_owner ? ctx->bases_owner->bases.p : ctx->bases.p);
}

// call before the first kernel that reads ctx->bases
static int wait_for_bases(msm_b200_ctx* ctx) {
  if (ctx->bases_pending) {
    CK(cudaStreamWaitEvent(ctx->stream, ctx->bases_ready, 0));
    ctx->bases_pending = false;
  }
  return 0;
}

// piece j of n scalars cut into `pieces` (multiples of 256 scalars, the last takes the rest)
static void scalar_piece(size_t n, int pieces, int j, size_t& lo, size_t& hi) {
  size_t per = ((n + pieces - 1) / pieces + 255) & ~(size_t)255;
  lo = std::min(n, per * (size_t)j);
  hi = (j == pieces - 1) ? n : std::min(n, lo + per);
}

// before the per-scalar kernels of piece j: wait for its upload (no-op for scalars that are already resident)
static int wait_for_scalars(msm_b200_ctx* ctx, int j) {
  if (ctx->sc_chunks > 0) CK(cudaStreamWaitEvent(ctx->stream, ctx->sc_ready[j], 0));
  return 0;
}

static int ceil_log2_sz(size_t n) {
  int k = 0;
  while (((size_t)1 << k) < n) k++;
  return k;
}

// CUDA-event stopwatch.  All Timer objects of one call share the context's event pool and its cursor
// (`ctx->ev_used`, reset by the C-ABI entry point), so marks taken at different levels never alias.
struct Timer {
  msm_b200_ctx* ctx;
  explicit Timer(msm_b200_ctx* c) : ctx(c) {}
  int mark(cudaStream_t on = nullptr) {  // records an event on the context's stream (or `on`), returns its index
    std::vector<cudaEvent_t>& ev = ctx->ev;
    if (ctx->ev_used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    cudaEventRecord(ev[ctx->ev_used], on ? on : ctx->stream);
    return (int)ctx->ev_used++;
  }
  float ms(int a, int b) {
    float t = 0;
    cudaEventElapsedTime(&t, ctx->ev[a], ctx->ev[b]);
    return t;
  }
};

static int resolve_timing(msm_b200_ctx* ctx, msm_b200_timing* tm) {
  const PendingTiming& pt = ctx->pending;
  tm->window_bits = pt.window_bits;
  tm->n_windows = pt.n_windows;
  tm->shared_buckets = pt.shared_buckets;
  if (!pt.valid) return 0;  // nothing was launched (empty input, all-zero scalars, digit dump)
  CK(cudaStreamSynchronize(ctx->stream));
  Timer T(ctx);
  tm->digits_ms = T.ms(pt.e[0], pt.e[1]);
  tm->sort_ms = T.ms(pt.e[1], pt.e[2]);
  tm->accumulate_ms = T.ms(pt.e[2], pt.e[3]);
  tm->reduce_ms = T.ms(pt.e[3], pt.e[4]);
  tm->hot_kernel_ms = 0;
  for (const auto& h : pt.hot) tm->hot_kernel_ms += T.ms(h.first, h.second);
  tm->hot_kernel_launches = (int)pt.hot.size();
  tm->window_bits = pt.window_bits;
  tm->n_windows = pt.n_windows;
  tm->rounds = pt.rounds;
  tm->n_adds = pt.n_adds;
  if (pt.h2d[0] >= 0) tm->h2d_ms = T.ms(pt.h2d[0], pt.h2d[1]);
  if (pt.fwd0[0] >= 0) {
    tm->fwd_round0_ms = T.ms(pt.fwd0[0], pt.fwd0[1]);
    tm->fwd_round0_pairs = (unsigned)std::min<unsigned long long>(pt.fwd0_pairs, 0xFFFFFFFFull);
  }
  tm->kernel_launches = ctx->launches;
  return 0;
}

// ------------------------------------------------------------------------------------------
// curve dispatch helpers
// ------------------------------------------------------------------------------------------
static bool field_is_large(int curve) { return curve == MSM_CURVE_BLS12_377_G1 || curve == MSM_CURVE_BLS12_381_G1; }
static int field_limbs(int curve) { return field_is_large(curve) ? 12 : 8; }
static int field_limbs29(int curve) { return field_is_large(curve) ? 14 : 9; }
static int field_bytes(int curve) { return field_is_large(curve) ? 48 : 32; }

static size_t point_bytes(int curve, int layout) {
  if (layout == MSM_LAYOUT_LE_BYTES) return 2 * (size_t)field_bytes(curve);
  if (curve == MSM_CURVE_ED_ON_BLS12_377) return 4 * 4 * (size_t)field_limbs29(curve);
  return 2 * 4 * (size_t)field_limbs29(curve) + 4;
}
static size_t scalar_bytes(int layout) { return layout == MSM_LAYOUT_LE_BYTES ? 32 : 36; }

// Engine default window size (the reference's table is tuned for 16 CPU threads,
// src/msm-common.ts:33-57; any c gives the same result).  The GPU wants windows that divide the
// scalar length evenly -- a short top window concentrates all its entries in a few buckets and
// costs extra tree rounds -- so: GLV halves (127/128 bits) use c = 16 (K = 8) from 2^14 points up
// (c = 19, K = 7 from 2^25: 12 % fewer additions outweigh the larger counting sort there),
// full-width scalars c = 14 (ed-on-bls12-377: 252 = 18 * 14) or 16 (254..256 bits).
static int default_window(int curve, int form, size_t n) {
  int lg = ceil_log2_sz(n);
  if (form == MSM_FORM_AFFINE_GLV) return lg >= 25 ? 19 : (lg >= 14 ? 16 : (lg >= 7 ? 8 : 4));
  if (curve == MSM_CURVE_ED_ON_BLS12_377) return lg >= 13 ? 14 : (lg >= 7 ? 9 : 4);  // 252 = 18 * 14 = 28 * 9
  return lg >= 13 ? 16 : (lg >= 7 ? 8 : 4);
}

// ------------------------------------------------------------------------------------------
// set_bases
// ------------------------------------------------------------------------------------------
// Window size for GLV bases WITH tables: all windows add into the same 2^(c-1) buckets, so the cost is
// 2n * ceil(128 / c) additions plus ONE bucket reduction over 2^(c-1) buckets.  Only window sizes that cut the
// window count matter: 16 (8 windows), 19 (7), 22 (6).  Measured crossovers (tools/sweep_tables.sh).
static int glv_table_window(const msm_b200_ctx* ctx, size_t n) {
  if (ctx->table_window > 0) return ctx->table_window;
  int lg = ceil_log2_sz(n);
  // BLS12-377: 2^21..2^23 -6 %, 2^24 -11 % against no tables; with 8-limb fields the scatter into 2^21 buckets
  // costs more than the sixth window saves
  // (Pallas: 19 only pays from 2^22 points: 2^21 6.22 (c = 16) / 6.26 ms (19), 2^23 21.6 / 21.1 ms)
  if (!field_is_large(ctx->curve)) return lg >= 22 ? 19 : 16;
  return lg >= 24 ? 22 : (lg >= 21 ? 19 : 16);
}

// `tables`: also build the window tables 2^(kc) G (resident bases only; the one-shot call passes false -- the
// tables cost about three MSMs to build)
template <class F, class G, uint32_t B3>
static int ingest_weierstrass(msm_b200_ctx* ctx, const void* d_in, size_t n, int layout, bool tables) {
  constexpr size_t REC = 2 * F::N * 4;  // bytes per record (x | y); two records per point
  ctx->table_c = ctx->table_K = 0;
  const int c = glv_table_window(ctx, n);
  const int K = (G::MAXBITS + 1 + c - 1) / c;
  tables = tables && ctx->tables_enabled && ceil_log2_sz(n) >= 14 && ceil_log2_sz(n) <= ctx->table_max_log2n &&
           (unsigned long long)2 * n * K < (1ull << 31);
  RET_IF(ensure(ctx, ctx->bases, n * 2 * REC * (tables ? K : 1)));
  LAUNCH(ctx, k_ingest_points<F>, cdiv(n, 128), 128, (const uint8_t*)d_in, n, layout, (uint4*)ctx->bases.p);
  if (tables) {
    for (int k = 1; k < K; k++)
      LAUNCH(ctx, (k_build_table<F, B3>), cdiv(n, 256), 256, (const uint4*)((const char*)ctx->bases.p + (size_t)(k - 1) * n * 2 * REC),
             (uint4*)((char*)ctx->bases.p + (size_t)k * n * 2 * REC), n, c);
    ctx->table_c = c;
    ctx->table_K = K;
  }
  CK(cudaGetLastError());
  return 0;
}

// Window size for twisted-Edwards bases WITH tables: every window adds into the same 2^(c-1) buckets, so the
// cost is n * ceil(252 / c) additions plus a bucket reduction over 2^(c-1) buckets only -- wider windows pay
// earlier than in the classic layout (c = 14, 18 bucket sets).
static int te_table_window(const msm_b200_ctx* ctx, size_t n) {
  if (ctx->table_window > 0) return ctx->table_window;
  int lg = ceil_log2_sz(n);
  return lg >= 19 ? 18 : (lg >= 17 ? 16 : 14);  // 2^22: 9.2 ms (c = 18) / 9.7 (20) / 10.3 (16) / 11.7 (no tables)
}

template <class F, class S>
static int ingest_te(msm_b200_ctx* ctx, const void* d_in, size_t n, int layout, bool tables) {
  constexpr size_t REC = 3 * F::N * 4;  // bytes per cached point (y+x | y-x | 2dxy)
  ctx->table_c = ctx->table_K = 0;
  const int c = te_table_window(ctx, n);
  const int K = (S::QBITS + 1 + c - 1) / c;
  tables = tables && ctx->tables_enabled && ceil_log2_sz(n) >= 13 && ceil_log2_sz(n) <= ctx->table_max_log2n &&
           (unsigned long long)n * K < (1ull << 31);
  RET_IF(ensure(ctx, ctx->bases, n * REC * (tables ? K : 1)));
  LAUNCH(ctx, k_te_ingest<F>, cdiv(n, 128), 128, (const uint8_t*)d_in, n, layout, (uint4*)ctx->bases.p);
  if (tables) {
    for (int k = 1; k < K; k++)
      LAUNCH(ctx, k_te_build_table<F>, cdiv(n, 256), 256, (const uint4*)((const char*)ctx->bases.p + (size_t)(k - 1) * n * REC),
             (uint4*)((char*)ctx->bases.p + (size_t)k * n * REC), n, c);
    ctx->table_c = c;
    ctx->table_K = K;
  }
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// batched inversion of the thread totals of one round (upper levels of the product tree)
// level 0 totals live in lvl_tot[0] (M1 elements); on return lvl_pre[0] holds their inverses.
// ------------------------------------------------------------------------------------------
template <class F>
static int invert_totals(msm_b200_ctx* ctx, size_t M1) {
  constexpr size_t FE = F::N * 4;
  size_t M[8];
  int ns = 0;  // serial levels
  M[0] = M1;
  // lvl_tot[l] = values of level l (M[l] elements); lvl_pre[l] = their prefixes / "others", then inverses
  while (M[ns] > (size_t)TREE_MAX && ns < 5) {
    size_t blocks = cdiv(M[ns], (size_t)UP_THREADS * UP_B1);
    M[ns + 1] = blocks * UP_THREADS;
    ns++;
  }
  const bool two = M[ns] > (size_t)TOP_CTA_MAX;  // scan level below the top block?
  const int top = ns + (two ? 1 : 0);
  if (two) M[top] = cdiv(M[ns], TREE_CTA);
  for (int l = 0; l <= top; l++) RET_IF(ensure(ctx, ctx->lvl_pre[l], (M[l] + 1) * FE));
  for (int l = 1; l <= top; l++) RET_IF(ensure(ctx, ctx->lvl_tot[l], (M[l] + 1) * FE));
  for (int l = 0; l < ns; l++)
    LAUNCH(ctx, k_up_fwd<F>, cdiv(M[l], (size_t)UP_THREADS * UP_B1), UP_THREADS, (const uint4*)ctx->lvl_tot[l].p, M[l],
           (uint4*)ctx->lvl_pre[l].p, (uint4*)ctx->lvl_tot[l + 1].p, M[l + 1]);
  if (two)
    LAUNCH(ctx, (k_tree_up<F, false, TREE_CTA>), (unsigned)M[top], TREE_CTA, (const uint4*)ctx->lvl_tot[ns].p, M[ns],
           (uint4*)ctx->lvl_pre[ns].p, (uint4*)ctx->lvl_tot[top].p, M[top]);
  if (M[top] <= (size_t)TREE_CTA)
    LAUNCH(ctx, (k_tree_top2<F, TREE_CTA / 2>), 1, TREE_CTA / 2, (const uint4*)ctx->lvl_tot[top].p, M[top],
           (uint4*)ctx->lvl_pre[top].p);
  else
    LAUNCH(ctx, (k_tree_top2<F, TREE_CTA>), 1, TREE_CTA, (const uint4*)ctx->lvl_tot[top].p, M[top],
           (uint4*)ctx->lvl_pre[top].p);
  if (two)
    LAUNCH(ctx, k_tree_down<F>, (unsigned)M[top], TREE_CTA, (uint4*)ctx->lvl_pre[ns].p, M[ns],
           (const uint4*)ctx->lvl_pre[top].p, M[top]);
  for (int l = ns - 1; l >= 0; l--)
    LAUNCH(ctx, k_up_bwd<F>, cdiv(M[l], (size_t)UP_THREADS * UP_B1), UP_THREADS, (const uint4*)ctx->lvl_tot[l].p, M[l],
           (uint4*)ctx->lvl_pre[l].p, (const uint4*)ctx->lvl_pre[l + 1].p, M[l + 1]);
  CK(cudaGetLastError());
  return 0;
}

template <class C>
static int zero_partial_t(msm_b200_ctx* ctx);

// pair-slot offsets of `rounds` tree rounds (ctx->cnt -> ctx->po, ctx->totals)
static int launch_scan(msm_b200_ctx* ctx, size_t NB, int rounds, uint32_t split = 0, int r0 = 0) {
  unsigned ntiles = cdiv(NB, SCAN_TILE);
  RET_IF(ensure(ctx, ctx->tilesum, (size_t)(MAX_ROUNDS + 1) * ntiles * 4));
  dim3 grid(ntiles, rounds);
  LAUNCH(ctx, k_scan_tiles, grid, SCAN_THREADS, (const uint32_t*)ctx->cnt.p, (uint32_t)NB, (uint32_t*)ctx->tilesum.p, ntiles,
         (unsigned long long*)ctx->totals.p, split, r0);
  LAUNCH(ctx, k_scan_write, grid, SCAN_THREADS, (const uint32_t*)ctx->cnt.p, (uint32_t)NB, (const uint32_t*)ctx->tilesum.p,
         ntiles, (uint32_t*)ctx->po.p, (unsigned long long*)ctx->totals.p, split, r0);
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// bucket reduction + Horner -> partial result in ctx->partial (any curve form)
// ------------------------------------------------------------------------------------------
template <class C, class Loader>
static int reduce_buckets(msm_b200_ctx* ctx, const Loader& ld, size_t NB, int K, int c) {
  constexpr size_t ITEM = (size_t)item_u4<C>() * 16;
  int remaining = c - 1;
  // level 0: 8 buckets per thread (measured best of 4 / 8 / 16 / 32 at 2^18 buckets: 1.59 / 1.28 / 1.44 / 1.94 ms
  // for the whole reduction); few buckets: 4 per lane quad
  const bool few = NB <= ((size_t)1 << 16);
  // (2^21 shared buckets, c = 22: 16 per thread -- 3.59 -> 3.10 ms; 32 per thread 2.91 ms but a longer serial tail)
  const int gb0 = ctx->reduce_gb0 > 0 ? ctx->reduce_gb0 : (few ? 2 : (NB >= ((size_t)1 << 20) ? 4 : 3));
  const size_t warp_min = ctx->reduce_warp_min > 0 ? ctx->reduce_warp_min : (few ? (size_t)1 << 16 : (size_t)4096);
  const bool quad0 = ctx->reduce_quad0 >= 0 ? ctx->reduce_quad0 != 0 : few;
  int gb = remaining < gb0 ? remaining : gb0;
  size_t items = NB >> gb;
  RET_IF(ensure(ctx, ctx->red[0], items * ITEM));
  RET_IF(ensure(ctx, ctx->red[1], (items / 2 + 1) * ITEM));
  if (quad0) LAUNCH(ctx, (k_reduce0_quad<C, Loader>), cdiv(items * 4, 64), 64, ld, (uint32_t)NB, gb, (uint4*)ctx->red[0].p);
  else LAUNCH(ctx, (k_reduce0<C, Loader>), cdiv(items, 64), 64, ld, (uint32_t)NB, gb, (uint4*)ctx->red[0].p);
  remaining -= gb;
  int cur = 0;
  while (remaining > 0) {
    if (items > warp_min) {
      gb = remaining < ctx->reduce_warp_gb ? remaining : ctx->reduce_warp_gb;  // one item per lane, groups of 2^gb lanes
      LAUNCH(ctx, (k_reduce_warp<C>), cdiv(items, 64), 64, (const uint4*)ctx->red[cur].p, (uint32_t)items, gb,
             (uint4*)ctx->red[cur ^ 1].p);
    } else {
      gb = remaining < 3 ? remaining : 3;  // latency-bound level: one item per lane quad
      LAUNCH(ctx, (k_reduce_quad<C>), cdiv(items * 4, 64), 64, (const uint4*)ctx->red[cur].p, (uint32_t)items, gb,
             (uint4*)ctx->red[cur ^ 1].p);
    }
    size_t out_items = items >> gb;
    items = out_items;
    remaining -= gb;
    cur ^= 1;
  }
  RET_IF(ensure(ctx, ctx->partial, 4 * 12 * 4));
  LAUNCH(ctx, (k_horner<C>), 1, 32, (const uint4*)ctx->red[cur].p, K, c, (uint4*)ctx->partial.p);
  CK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// generic bucket method (msmBasic): twisted Edwards and msmProjective
// ------------------------------------------------------------------------------------------
template <class C, class S>
static int run_bucket_basic(msm_b200_ctx* ctx, const void* d_scalars, size_t n, int layout, int c,
                            msm_b200_timing* tm) {
  using F = typename C::F;
  constexpr size_t FE = F::N * 4;
  Timer T(ctx);
  ctx->pending.valid = false;
  const int b = S::QBITS;  // Scalar.sizeInBits, src/msm-basic.ts:55
  const int K = (b + 1 + c - 1) / c;
  ctx->pending.window_bits = c;
  ctx->pending.n_windows = K;
  const uint32_t L = 1u << (c - 1);
  // shared buckets: the resident bases carry one table 2^(kc) P per window (twisted Edwards: k_te_build_table);
  // the Weierstrass tables are laid out for the GLV half scalars and only match here if K does
  const bool shared = C::BASE_STRIDE == 1 && ctx->table_c == c && ctx->table_K == K;
  const int KR = shared ? 1 : K;
  const size_t NB = (size_t)KR * L;
  ctx->pending.shared_buckets = shared ? 1 : 0;
  if (NB > ((size_t)1 << 28)) return fail(ctx, MSM_E_INVALID, "window too large");
  // sorted-entry slots are addressed with 32 bits (2 * pair offset + position)
  if ((unsigned long long)n * K >= (1ull << 31))
    return fail(ctx, MSM_E_INVALID, "too many digit entries for one context (n * windows >= 2^31): shard the points");
  int e0 = T.mark();
  RET_IF(ensure(ctx, ctx->hs, n * 32));
  RET_IF(ensure(ctx, ctx->cnt, NB * 4));
  RET_IF(ensure(ctx, ctx->cursor, NB * 4));
  RET_IF(ensure(ctx, ctx->po, 2 * NB * 4));
  RET_IF(ensure(ctx, ctx->totals, N_TOTALS * 8));
  CK(cudaMemsetAsync(ctx->cnt.p, 0, NB * 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->cursor.p, 0, NB * 4, ctx->stream));
  SortArgs sa;
  sa.hs = (const uint4*)ctx->hs.p;
  sa.S = n;
  sa.c = c;
  sa.K = K;
  sa.L = L;
  sa.cnt = (uint32_t*)ctx->cnt.p;
  sa.cursor = (uint32_t*)ctx->cursor.p;
  sa.po0 = (const uint32_t*)ctx->po.p;
  sa.ent = nullptr;
  sa.pairkey = nullptr;
  sa.digits = nullptr;
  sa.bucket_stride = shared ? 0u : L;
  sa.cnt_stride = sa.bucket_stride;
  sa.ent_stride = shared ? (uint32_t)ctx->n_bases : 0u;
  {  // unpack + count, piece by piece behind the scalar upload
    const int pieces = ctx->sc_chunks > 0 ? ctx->sc_chunks : 1;
    const size_t sb = scalar_bytes(layout);
    for (int j = 0; j < pieces; j++) {
      size_t lo, hi;
      scalar_piece(n, pieces, j, lo, hi);
      RET_IF(wait_for_scalars(ctx, j));
      if (hi <= lo) continue;
      LAUNCH(ctx, k_load_scalars<S>, cdiv(hi - lo, 128), 128, (const uint8_t*)d_scalars + lo * sb, hi - lo, layout,
             (uint4*)ctx->hs.p + 2 * lo);
      SortArgs sp = sa;
      sp.hs = sa.hs + 2 * lo;
      sp.S = hi - lo;
      LAUNCH(ctx, k_hist_scatter8<false>, cdiv(hi - lo, 256), 256, sp);
    }
    ctx->sc_chunks = 0;
  }
  int e1 = T.mark();
  CK(cudaMemsetAsync(ctx->totals.p, 0, N_TOTALS * 8, ctx->stream));
  // entries per virtual bucket: 128 when there are plenty of entries; fewer for smaller inputs, so that the accumulation
  // still runs on ~2^19 threads (each thread is one serial chain of additions)
  // (measured, tools/perf_sweep.py with MSM_B200_BUCKET_SPLIT: ed-on-bls12-377 2^16 0.73 -> 0.54 ms with 16, 2^18 1.34 ->
  // 1.04 ms with 32, 2^22 9.0 -> 8.5 ms with 64, 2^24 best with 128; the 12-limb projective form, few warps per SM at
  // 178 registers, wants more threads still: msmProjective 2^22 43.0 -> 36.2 ms with 32)
  uint32_t vsplit = ctx->bucket_split > 0 ? (uint32_t)ctx->bucket_split : (uint32_t)BUCKET_SPLIT;
  if (ctx->bucket_split <= 0) {
    const unsigned long long entries = (unsigned long long)n * K;
    while (vsplit > 32 && entries / vsplit < (1ull << 19)) vsplit >>= 1;
    if (F::N > 8 && vsplit > 32) vsplit = 32;
    if (entries < (1ull << 21)) vsplit = 16;
  }
  RET_IF(launch_scan(ctx, NB, 2, vsplit));  // po[0]: padded entry offsets, po[1]: virtual-bucket offsets
  CK(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, N_TOTALS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  const size_t P0 = ctx->h_totals[0];
  RET_IF(ensure(ctx, ctx->ent, (2 * P0 + 2) * 4));
  sa.ent = (uint32_t*)ctx->ent.p;
  LAUNCH(ctx, k_hist_scatter8<true>, cdiv(n, 256), 256, sa);
  int e2 = T.mark();
  RET_IF(ensure(ctx, ctx->buckets, NB * C::ACC_FE * FE));
  RET_IF(wait_for_bases(ctx));
  const size_t V = ctx->h_totals[1];
  const bool split = ctx->h_totals[MAX_ROUNDS + 1] > (unsigned long long)vsplit;
  if (split) {
    RET_IF(ensure(ctx, ctx->pairkey[0], (V + 1) * 4));
    RET_IF(ensure(ctx, ctx->elem[0], (V + 1) * C::ACC_FE * FE));
    LAUNCH(ctx, k_fill_pairkey, cdiv((NB + 31) / 32 * 32, 256), 256, (const uint32_t*)ctx->po.p + NB, (uint32_t)NB,
           (const unsigned long long*)ctx->totals.p + 1 /* = V */, (uint32_t*)ctx->pairkey[0].p);
  }
  int h0 = T.mark();
  if (split) {
    LAUNCH(ctx, k_bucket_acc_v<C>, cdiv(V, 128), 128, (const uint32_t*)ctx->cnt.p, (const uint32_t*)ctx->po.p,
           (const uint32_t*)ctx->po.p + NB, (const uint32_t*)ctx->pairkey[0].p, (const uint32_t*)ctx->ent.p,
           bases_ptr(ctx), (uint32_t)V, vsplit, (uint4*)ctx->elem[0].p);
    // buckets with many pieces: halving passes; with few: a short serial loop in k_bucket_combine
    const unsigned long long max_pieces = (ctx->h_totals[MAX_ROUNDS + 1] + vsplit - 1) / vsplit;
    int serial_max = 1 << 30;
    if (max_pieces > 16) {
      serial_max = 0;  // every multi-piece bucket goes through the tree
      for (int p = 0; (1ull << p) < max_pieces; p++)
        LAUNCH(ctx, k_bucket_tree_pass<C>, cdiv(V, 128), 128, (const uint32_t*)ctx->po.p + NB,
               (const uint32_t*)ctx->pairkey[0].p, (uint4*)ctx->elem[0].p, (uint32_t)V, (uint32_t)V, p);
    }
    LAUNCH(ctx, k_bucket_combine<C>, cdiv(NB, 128), 128, (const uint32_t*)ctx->cnt.p, (const uint32_t*)ctx->po.p + NB,
           (const uint4*)ctx->elem[0].p, (uint32_t)NB, vsplit, serial_max, (uint4*)ctx->buckets.p);
  } else {
    LAUNCH(ctx, k_bucket_acc<C>, cdiv(NB, 128), 128, (const uint32_t*)ctx->cnt.p, (const uint32_t*)ctx->po.p,
           (const uint32_t*)ctx->ent.p, bases_ptr(ctx), (uint32_t)NB, (uint4*)ctx->buckets.p);
  }
  int h1 = T.mark();
  CK(cudaGetLastError());
  int e3 = T.mark();
  AccBucketLoader<C> ld;
  ld.buckets = (const uint4*)ctx->buckets.p;
  RET_IF((reduce_buckets<C>(ctx, ld, NB, KR, c)));
  int e4 = T.mark();
  {
    PendingTiming& pt = ctx->pending;
    pt.valid = true;
    pt.e[0] = e0, pt.e[1] = e1, pt.e[2] = e2, pt.e[3] = e3, pt.e[4] = e4;
    pt.hot.assign(1, std::make_pair(h0, h1));
    pt.window_bits = c;
    pt.n_windows = K;
    pt.rounds = 0;
    pt.n_adds = ctx->h_totals[MAX_ROUNDS + 2];  // one mixed addition per sorted entry
  }
  if (tm) RET_IF(resolve_timing(ctx, tm));
  return 0;
}

template <class C>
static int finalize_any(msm_b200_ctx* ctx, const void* partials_dev, int count, msm_b200_point* out) {
  using F = typename C::F;
  RET_IF(ensure(ctx, ctx->result, (2 * F::N + 1) * 4));
  LAUNCH(ctx, (k_finalize<C>), 1, 32, (const uint4*)partials_dev, count, (uint32_t*)ctx->result.p);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(ctx->h_result, ctx->result.p, (2 * F::N + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  memset(out, 0, sizeof *out);
  memcpy(out->x, ctx->h_result, F::N * 4);
  memcpy(out->y, ctx->h_result + F::N, F::N * 4);
  out->is_zero = (int32_t)ctx->h_result[2 * F::N];
  return 0;
}

template <class C, class S>
static int random_points_t(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed) {
  using F = typename C::F;
  size_t entries = (size_t)RP_TABLES << RP_BITS;
  RET_IF(ensure(ctx, ctx->rp_tables, entries * 2 * F::N * 4));
  LAUNCH(ctx, k_rp_tables<C>, cdiv(entries, 64), 64, (uint4*)ctx->rp_tables.p, seed);
  size_t threads = (n + RP_BATCH - 1) / RP_BATCH;
  LAUNCH(ctx, k_rp_points<C>, cdiv(threads, 64), 64, (const uint4*)ctx->rp_tables.p, (uint8_t*)dst_dev, n,
         seed ^ 0x5EEDull, first);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Weierstrass GLV batched-affine MSM -> projective partial in ctx->partial
// ------------------------------------------------------------------------------------------
template <class F, class G, uint32_t B3>
static int run_affine_glv(msm_b200_ctx* ctx, const void* d_scalars, size_t n, int layout, int c,
                          msm_b200_timing* tm, uint32_t* digits_dump_dev) {
  constexpr size_t FE = F::N * 4;
  Timer T(ctx);
  ctx->pending.valid = false;
  ctx->pending.fwd0[0] = ctx->pending.fwd0[1] = -1;
  const int b = G::MAXBITS;  // Scalar.maxBits, src/wasm/glv.ts:216-226
  const int K = (b + 1 + c - 1) / c;
  ctx->pending.window_bits = c;
  ctx->pending.n_windows = K;
  const uint32_t L = 1u << (c - 1);
  // shared buckets: the resident bases carry a table 2^(kc) G per window, so all windows add into ONE set
  // of L buckets (no Horner step, K times fewer buckets to reduce); the digit dump of the tests keeps the
  // classic layout
  const bool shared = ctx->table_c == c && ctx->table_K == K && !digits_dump_dev;
  const int KR = shared ? 1 : K;  // bucket sets to reduce
  const size_t NB = (size_t)KR * L;
  const size_t S = 2 * n;
  ctx->pending.shared_buckets = shared ? 1 : 0;
  if (NB > ((size_t)1 << 28)) return fail(ctx, MSM_E_INVALID, "window too large");
  // sorted-entry slots are addressed with 32 bits (2 * pair offset + position)
  if ((unsigned long long)S * K >= (1ull << 31))
    return fail(ctx, MSM_E_INVALID, "too many digit entries for one context (2n * windows >= 2^31): shard the points");

  int e0 = T.mark();
  // --- GLV + digits + histogram
  // shared buckets with few buckets: count per (window, bucket) and merge (see SortArgs::cnt)
  const bool split_counts = shared && L <= (1u << 16);
  const size_t NBK = (shared && !split_counts) ? (size_t)L : (size_t)K * L;  // counters the two sort passes touch
  RET_IF(ensure(ctx, ctx->hs, S * 16));
  RET_IF(ensure(ctx, ctx->cnt, NB * 4));
  RET_IF(ensure(ctx, ctx->cursor, NBK * 4));
  if (split_counts) RET_IF(ensure(ctx, ctx->cntk, NBK * 4));
  RET_IF(ensure(ctx, ctx->po, (size_t)(MAX_ROUNDS + 1) * NB * 4));
  RET_IF(ensure(ctx, ctx->totals, N_TOTALS * 8));
  uint32_t* cnt_hist = split_counts ? (uint32_t*)ctx->cntk.p : (uint32_t*)ctx->cnt.p;
  CK(cudaMemsetAsync(cnt_hist, 0, NBK * 4, ctx->stream));
  if (!split_counts) CK(cudaMemsetAsync(ctx->cursor.p, 0, NBK * 4, ctx->stream));
  SortArgs sa;
  sa.hs = (const uint4*)ctx->hs.p;
  sa.S = S;
  sa.c = c;
  sa.K = K;
  sa.L = L;
  sa.cnt = cnt_hist;
  sa.cursor = (uint32_t*)ctx->cursor.p;
  sa.po0 = (const uint32_t*)ctx->po.p;
  sa.ent = nullptr;
  sa.pairkey = nullptr;
  sa.digits = digits_dump_dev;
  sa.bucket_stride = shared ? 0u : L;
  sa.cnt_stride = (shared && !split_counts) ? 0u : L;
  sa.ent_stride = shared ? (uint32_t)(2 * ctx->n_bases) : 0u;
  {  // decompose + count, piece by piece behind the scalar upload (the digit dump of the tests: one piece)
    const int pieces = (ctx->sc_chunks > 0 && !digits_dump_dev) ? ctx->sc_chunks : 1;
    const size_t sb = scalar_bytes(layout);
    for (int j = 0; j < pieces; j++) {
      size_t lo, hi;
      scalar_piece(n, pieces, j, lo, hi);
      if (ctx->sc_chunks > 0) {
        if (pieces > 1) RET_IF(wait_for_scalars(ctx, j));
        else for (int u = 0; u < ctx->sc_chunks; u++) RET_IF(wait_for_scalars(ctx, u));
      }
      if (hi <= lo) continue;
      LAUNCH(ctx, k_glv<G>, cdiv(hi - lo, 128), 128, (const uint8_t*)d_scalars + lo * sb, hi - lo, layout, (uint4*)ctx->hs.p + 2 * lo);
      SortArgs sp = sa;
      sp.hs = sa.hs + 2 * lo;
      sp.S = 2 * (hi - lo);
      LAUNCH(ctx, k_hist_scatter<false>, cdiv(2 * (hi - lo), 256), 256, sp);
    }
    ctx->sc_chunks = 0;
  }
  if (split_counts)
    LAUNCH(ctx, k_merge_counts, cdiv(L, 256), 256, (const uint32_t*)ctx->cntk.p, K, L, (uint32_t*)ctx->cnt.p,
           (uint32_t*)ctx->cursor.p);
  int e1 = T.mark();
  // --- offsets for every round; the host needs the totals (one sync), the scatter does not: it is queued
  //     first, sized by the upper bound 2 * P0 <= S * K + NB, so the GPU keeps working while the host wakes up
  CK(cudaMemsetAsync(ctx->totals.p, 0, N_TOTALS * 8, ctx->stream));
  RET_IF(launch_scan(ctx, NB, SCAN_ROUNDS_FIRST));  // rounds 0 .. SCAN_ROUNDS_FIRST - 1; more below if a bucket is larger
  CK(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, N_TOTALS * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaEventRecord(ctx->totals_ready, ctx->stream));
  if (!digits_dump_dev) {
    const size_t slots_max = S * (size_t)K + NB + 2;
    RET_IF(ensure(ctx, ctx->ent, slots_max * 4));
    RET_IF(ensure(ctx, ctx->pairkey[0], (slots_max / 2 + 2) * 4));
    sa.ent = (uint32_t*)ctx->ent.p;
    sa.pairkey = (uint32_t*)ctx->pairkey[0].p;
    CK(cudaMemsetAsync(ctx->ent.p, 0, slots_max * 4, ctx->stream));  // padding slots are read (and discarded)
    LAUNCH(ctx, k_hist_scatter<true>, cdiv(S, 256), 256, sa);
    LAUNCH(ctx, k_fill_pairkey, cdiv((NB + 31) / 32 * 32, 256), 256, (const uint32_t*)ctx->po.p, (uint32_t)NB,
           (const unsigned long long*)ctx->totals.p, (uint32_t*)ctx->pairkey[0].p);
  }
  CK(cudaEventSynchronize(ctx->totals_ready));
  const unsigned long long maxcnt = ctx->h_totals[MAX_ROUNDS + 1];
  int R = 1;
  while (((unsigned long long)1 << R) < maxcnt) R++;
  if (R > MAX_ROUNDS - 1) return fail(ctx, MSM_E_INVALID, "bucket too large");
  if (R + 1 > SCAN_ROUNDS_FIRST && !digits_dump_dev) {  // large buckets: offsets of the remaining rounds
    RET_IF(launch_scan(ctx, NB, R + 1 - SCAN_ROUNDS_FIRST, 0, SCAN_ROUNDS_FIRST));
    CK(cudaMemcpyAsync(ctx->h_totals, ctx->totals.p, N_TOTALS * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->totals_ready, ctx->stream));
    CK(cudaEventSynchronize(ctx->totals_ready));
  }
  const size_t P0 = ctx->h_totals[0];
  if (digits_dump_dev) {  // tests only need the digits
    CK(cudaStreamSynchronize(ctx->stream));
    if (tm) {
      tm->window_bits = c;
      tm->n_windows = K;
    }
    return 0;
  }
  unsigned long long n_adds = 0;
  if (P0 == 0) {  // every digit is zero (e.g. all scalars are 0): the sum is the neutral element
    RET_IF((zero_partial_t<WeierCurve<F, B3>>(ctx)));
    if (tm) {
      tm->window_bits = c;
      tm->n_windows = K;
    }
    return 0;
  }
  RET_IF(ensure(ctx, ctx->pairkey[1], (ctx->h_totals[1] + 1) * 4));
  int e2 = T.mark();
  // --- tree rounds
  RET_IF(ensure(ctx, ctx->elem[0], ElemBuf<F>::bytes(ctx->h_totals[1] + 1)));
  RET_IF(ensure(ctx, ctx->elem[1], ElemBuf<F>::bytes(ctx->h_totals[2] + 1)));
  RET_IF(ensure(ctx, ctx->prefix, (P0 + 1) * FE));
  RET_IF(ensure(ctx, ctx->fin, FinBuf<F>::bytes(NB)));
  std::vector<std::pair<int, int>> hot;
  int rounds_run = 0;
  bool finished_projective = false;
  RET_IF(wait_for_bases(ctx));
  for (int r = 0; r < R; r++) {
    const size_t P = ctx->h_totals[r];
    const size_t Pn = ctx->h_totals[r + 1];
    if (P == 0) break;
    RoundArgs<F> a;
    a.r = r;
    a.P = P;
    a.cnt = (const uint32_t*)ctx->cnt.p;
    a.po_r = (const uint32_t*)ctx->po.p + (size_t)r * NB;
    a.po_n = (const uint32_t*)ctx->po.p + (size_t)(r + 1) * NB;
    a.pairkey = (const uint32_t*)ctx->pairkey[r & 1].p;
    a.pairkey_next = (uint32_t*)ctx->pairkey[(r + 1) & 1].p;
    a.ent = (const uint32_t*)ctx->ent.p;
    a.bases = bases_ptr(ctx);
    // elements of round r (r >= 1) live in elem[(r-1)&1] with capacity P_r; outputs go to elem[r&1]
    a.in.base = (uint4*)ctx->elem[(r + 1) & 1].p;
    a.in.cap = P;
    a.out.base = (uint4*)ctx->elem[r & 1].p;
    a.out.cap = Pn;
    a.fin.base = (uint4*)ctx->fin.p;
    a.fin.cap = NB;
    a.prefix = nullptr;
    a.tot = nullptr;
    a.invtot = nullptr;
    a.M1 = 0;
    a.B0 = 0;
    a.blktot = nullptr;
    // tail: one thread per bucket finishes the sums with projective mixed additions once that is cheaper
    // than the rounds still to come.  Cost model in units of field products at full rate, fitted with
    // tools/sweep_finish.sh: a round costs finish_round_modmuls (product tree, inversion, launch gaps) on
    // top of 6 per addition; the tail costs finish_add_modmuls per addition (11 products, times the
    // divergence of a thread-per-bucket loop).  Only with few elements per bucket (serial chain).
    unsigned long long adds_left = 0;
    for (int s = r; s < R; s++) adds_left += ctx->h_totals[MAX_ROUNDS + 3 + s];
    const int fin_elems = ctx->finish_max_elems > 0 ? ctx->finish_max_elems : (shared ? 2 * FINISH_MAX_ELEMS : FINISH_MAX_ELEMS);
    const double fin_round = ctx->finish_round_modmuls > 0 ? ctx->finish_round_modmuls : (shared ? 3.0e6 : 1.0e6);
    if (((maxcnt + (1ull << r) - 1) >> r) <= (unsigned long long)fin_elems &&
        (double)adds_left * (ctx->finish_add_modmuls - 6.0) <= (double)(R - r) * fin_round) {
      RET_IF(ensure(ctx, ctx->buckets, NB * 3 * FE));
      if (2 * P >= 3 * NB) {  // dense buckets (3+ elements each on average): one thread per bucket
        if (r == 0) LAUNCH(ctx, (k_finish_buckets<F, B3, true>), cdiv(2 * NB, 64), 64, a, (uint32_t)NB, (uint4*)ctx->buckets.p);
        else LAUNCH(ctx, (k_finish_buckets<F, B3, false>), cdiv(2 * NB, 64), 64, a, (uint32_t)NB, (uint4*)ctx->buckets.p);
      } else if (r == 0) {
        LAUNCH(ctx, (k_finish_slots<F, B3, true>), cdiv(P, 64), 64, a, (uint4*)ctx->buckets.p);
        LAUNCH(ctx, (k_finish_rest<F, true>), cdiv(NB, 128), 128, a, (uint32_t)NB, (uint4*)ctx->buckets.p);
      } else {
        LAUNCH(ctx, (k_finish_slots<F, B3, false>), cdiv(P, 64), 64, a, (uint4*)ctx->buckets.p);
        LAUNCH(ctx, (k_finish_rest<F, false>), cdiv(NB, 128), 128, a, (uint32_t)NB, (uint4*)ctx->buckets.p);
      }
      finished_projective = true;
      break;
    }
    // pairs per thread: small rounds run as ONE full wave of resident blocks (no tail, few thread
    // totals for the product tree); large rounds as ~ACC_WAVES waves so that dynamic block
    // scheduling evens out the SMs
    size_t per_wave = (size_t)acc_resident<F>() * ctx->sm_count * ACC_THREADS;
    int B0;
    if (P <= (size_t)ACC_SINGLE_WAVE_MAX) {
      B0 = (int)cdiv(P, per_wave);
      if (B0 < ctx->acc_min_pairs) B0 = ctx->acc_min_pairs;
    } else {
      B0 = (int)cdiv(P, per_wave * ACC_WAVES);
      if (B0 > ACC_MAX_PAIRS) B0 = ACC_MAX_PAIRS;
    }
    a.B0 = B0;
    unsigned grid = cdiv(P, (size_t)ACC_THREADS * B0);
    size_t M1 = (size_t)grid * ACC_THREADS;
    const bool blk = grid <= (unsigned)TOP_CTA_MAX;  // block totals go straight to the top block
    a.prefix = (uint4*)ctx->prefix.p;
    a.M1 = M1;
    size_t Mtree;
    if (blk) {
      RET_IF(ensure(ctx, ctx->others, M1 * FE));
      RET_IF(ensure(ctx, ctx->lvl_tot[0], ((size_t)grid + 1) * FE));
      RET_IF(ensure(ctx, ctx->lvl_pre[0], ((size_t)grid + 1) * FE));
      a.tot = (uint4*)ctx->others.p;
      a.blktot = (uint4*)ctx->lvl_tot[0].p;
      Mtree = grid;
    } else {
      RET_IF(ensure(ctx, ctx->lvl_tot[0], (M1 + 1) * FE));
      RET_IF(ensure(ctx, ctx->lvl_pre[0], (M1 + 1) * FE));
      a.tot = (uint4*)ctx->lvl_tot[0].p;
      a.blktot = nullptr;
      Mtree = M1;
    }
    a.invtot = nullptr;  // set after invert_totals (which may grow its buffers)
    if (r == 0) {  // the round-0 forward pass is the HBM-bound kernel of the path (random gathers): timed on its own
      int f0 = T.mark();
      if (blk) LAUNCH(ctx, (k_fwd<F, true, true>), grid, ACC_THREADS, a);
      else LAUNCH(ctx, (k_fwd<F, true, false>), grid, ACC_THREADS, a);
      int f1 = T.mark();
      ctx->pending.fwd0[0] = f0, ctx->pending.fwd0[1] = f1;
      ctx->pending.fwd0_pairs = P;
    } else {
      if (blk) LAUNCH(ctx, (k_fwd<F, false, true>), grid, ACC_THREADS, a);
      else LAUNCH(ctx, (k_fwd<F, false, false>), grid, ACC_THREADS, a);
    }
    RET_IF(invert_totals<F>(ctx, Mtree));
    a.invtot = (const uint4*)ctx->lvl_pre[0].p;
    int h0 = T.mark();
    if (r == 0) {
      if (blk) LAUNCH(ctx, (k_bwd<F, true, true>), grid, ACC_THREADS, a);
      else LAUNCH(ctx, (k_bwd<F, true, false>), grid, ACC_THREADS, a);
    } else {
      if (blk) LAUNCH(ctx, (k_bwd<F, false, true>), grid, ACC_THREADS, a);
      else LAUNCH(ctx, (k_bwd<F, false, false>), grid, ACC_THREADS, a);
    }
    int h1 = T.mark();
    hot.push_back({h0, h1});
    n_adds += ctx->h_totals[MAX_ROUNDS + 3 + r];  // additions finished by this k_bwd launch
    rounds_run++;
  }
  CK(cudaGetLastError());
  int e3 = T.mark();
  // --- bucket reduction
  if (finished_projective) {
    AccBucketLoader<WeierCurve<F, B3>> ld;
    ld.buckets = (const uint4*)ctx->buckets.p;
    RET_IF((reduce_buckets<WeierCurve<F, B3>>(ctx, ld, NB, KR, c)));
  } else {
    AffineBucketLoader<F, B3> ld;
    ld.fin = (const uint4*)ctx->fin.p;
    ld.cap = NB;
    ld.cnt = (const uint32_t*)ctx->cnt.p;
    RET_IF((reduce_buckets<WeierCurve<F, B3>>(ctx, ld, NB, KR, c)));
  }
  int e4 = T.mark();
  {
    PendingTiming& pt = ctx->pending;
    pt.valid = true;
    pt.e[0] = e0, pt.e[1] = e1, pt.e[2] = e2, pt.e[3] = e3, pt.e[4] = e4;
    pt.hot = hot;
    pt.window_bits = c;
    pt.n_windows = K;
    pt.rounds = rounds_run;
    pt.n_adds = n_adds;
  }
  if (tm) RET_IF(resolve_timing(ctx, tm));
  return 0;
}


// ------------------------------------------------------------------------------------------
// per-curve entry points (one translation unit each, so the curves build in parallel)
// ------------------------------------------------------------------------------------------
struct CurveOps {
  int (*ingest)(msm_b200_ctx*, const void* d_in, size_t n, int layout, bool tables);
  int (*run)(msm_b200_ctx*, const void* d_scalars, size_t n, int layout, int form, int c, msm_b200_timing*,
             uint32_t* digits_dump_dev);
  int (*zero_partial)(msm_b200_ctx*);
  int (*finalize)(msm_b200_ctx*, const void* partials_dev, int count, msm_b200_point* out);
  int (*random_points)(msm_b200_ctx*, void* dst_dev, size_t first, size_t n, uint64_t seed);
  int (*random_scalars)(msm_b200_ctx*, void* dst_dev, size_t first, size_t n, uint64_t seed);
};
const CurveOps* curve_ops_bls377();
const CurveOps* curve_ops_pallas();
const CurveOps* curve_ops_ed377();
const CurveOps* curve_ops_bls381();

template <class C>
static int zero_partial_t(msm_b200_ctx* ctx) {
  RET_IF(ensure(ctx, ctx->partial, 4 * 12 * 4));
  LAUNCH(ctx, (k_zero_partial<C>), 1, 32, (uint4*)ctx->partial.p);
  CK(cudaGetLastError());
  return 0;
}

template <class S>
static int random_scalars_t(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed) {
  LAUNCH(ctx, k_random_scalars<S>, cdiv(n, 256), 256, (uint32_t*)dst_dev, n, seed, first);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// Weierstrass curve with GLV: forms AFFINE_GLV and PROJECTIVE
template <class F, class G, uint32_t B3>
static int run_weierstrass_t(msm_b200_ctx* ctx, const void* d_s, size_t n, int layout, int form, int c,
                             msm_b200_timing* tm, uint32_t* digits_dump_dev) {
  if (form == MSM_FORM_AFFINE_GLV) return run_affine_glv<F, G, B3>(ctx, d_s, n, layout, c, tm, digits_dump_dev);
  return run_bucket_basic<WeierCurve<F, B3>, G>(ctx, d_s, n, layout, c, tm);
}

#define MSM_DEFINE_WEIERSTRASS_CURVE(fn, F, G, B3)                                                          \
  const CurveOps* fn() {                                                                                    \
    static const CurveOps ops = {ingest_weierstrass<F, G, B3>,   run_weierstrass_t<F, G, B3>,                \
                                 zero_partial_t<WeierCurve<F, B3>>, finalize_any<WeierCurve<F, B3>>,         \
                                 random_points_t<WeierCurve<F, B3>, G>, random_scalars_t<G>};                \
    return &ops;                                                                                            \
  }
