// libmsm_b200.so -- the C ABI (include/msm_b200.h) over the per-curve engines.
#include <cstdlib>
#include <algorithm>
#include "../../include/msm_b200_test.h"
#include "engine.cuh"
#include "microbench.cuh"
#include "inv_quad.cuh"

std::string& msm_global_err() {
  static thread_local std::string e;
  return e;
}

static const CurveOps* ops_of(int curve) {
  switch (curve) {
    case MSM_CURVE_BLS12_377_G1: return curve_ops_bls377();
    case MSM_CURVE_PALLAS: return curve_ops_pallas();
    case MSM_CURVE_ED_ON_BLS12_377: return curve_ops_ed377();
    case MSM_CURVE_BLS12_381_G1: return curve_ops_bls381();
  }
  return nullptr;
}

// ends a loan of resident bases (msm_b200_share_bases), if there is one
static void drop_borrowed_bases(msm_b200_ctx* ctx) {
  if (!ctx->bases_owner) return;
  std::vector<msm_b200_ctx*>& v = ctx->bases_owner->borrowers;
  v.erase(std::remove(v.begin(), v.end(), ctx), v.end());
  ctx->bases_owner = nullptr;
}

static int set_bases_impl(msm_b200_ctx* ctx, const void* points, size_t n, int layout, int on_device,
                          bool overlapped = false) {
  if (!ctx) return fail(nullptr, MSM_E_INVALID, "null context");
  if (ctx->bases_pending) {  // a previous overlapped upload that was never consumed
    CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->bases_pending = false;
  }
  if (layout != MSM_LAYOUT_LIMB29_MONT && layout != MSM_LAYOUT_LE_BYTES) return fail(ctx, MSM_E_INVALID, "bad point layout");
  if (n == 0 || !points) {
    ctx->n_bases = 0;
    return n == 0 ? 0 : fail(ctx, MSM_E_INVALID, "null points");
  }
  if (n > ((size_t)1 << 30)) return fail(ctx, MSM_E_INVALID, "too many points (max 2^30)");
  CK(cudaSetDevice(ctx->device));
  drop_borrowed_bases(ctx);    // own bases from now on
  ctx->bases_gen++;            // contexts that borrowed the previous set must not read the new one by accident
  const void* d_in = points;
  cudaStream_t main_stream = ctx->stream;
  if (overlapped) {
    // everything queued on the main stream so far (earlier calls reading the old bases) first
    CK(cudaEventRecord(ctx->bases_ready, main_stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->bases_ready, 0));
    ctx->stream = ctx->copy_stream;
  }
  int rc = 0;
  if (!on_device) {
    size_t bytes = n * point_bytes(ctx->curve, layout);
    rc = ensure(ctx, ctx->raw_points, bytes);
    if (rc == 0 && cudaMemcpyAsync(ctx->raw_points.p, points, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
      rc = fail(ctx, MSM_E_CUDA, "H2D copy of the points failed");
    d_in = ctx->raw_points.p;
  }
  if (rc == 0) rc = ops_of(ctx->curve)->ingest(ctx, d_in, n, layout, /*tables=*/!overlapped);
  if (overlapped) {
    if (rc == 0 && cudaEventRecord(ctx->bases_ready, ctx->copy_stream) == cudaSuccess) ctx->bases_pending = true;
    ctx->stream = main_stream;
  }
  RET_IF(rc);
  ctx->n_bases = n;
  return 0;
}

// ------------------------------------------------------------------------------------------
// run: scalars -> partial (device)
// ------------------------------------------------------------------------------------------
static int run_partial_impl(msm_b200_ctx* ctx, const void* scalars, size_t n, int layout, int on_device, int form,
                            int window_bits, msm_b200_timing* tm, uint32_t* digits_dump_dev = nullptr) {
  if (!ctx) return fail(nullptr, MSM_E_INVALID, "null context");
  if (layout != MSM_LAYOUT_LIMB29_MONT && layout != MSM_LAYOUT_LE_BYTES) return fail(ctx, MSM_E_INVALID, "bad scalar layout");
  if (n > ctx->n_bases) return fail(ctx, MSM_E_STATE, "more scalars than resident bases (call set_bases first)");
  if (ctx->bases_owner && ctx->bases_owner->bases_gen != ctx->borrowed_gen)
    return fail(ctx, MSM_E_STATE, "the shared bases were replaced by their owner (call msm_b200_share_bases again)");
  if (n > 0 && !scalars) return fail(ctx, MSM_E_INVALID, "null scalars");
  const bool te = ctx->curve == MSM_CURVE_ED_ON_BLS12_377;
  if (te && form != MSM_FORM_TE_EXTENDED) return fail(ctx, MSM_E_INVALID, "twisted Edwards curve needs MSM_FORM_TE_EXTENDED");
  if (!te && form != MSM_FORM_AFFINE_GLV && form != MSM_FORM_PROJECTIVE)
    return fail(ctx, MSM_E_INVALID, "Weierstrass curve needs MSM_FORM_AFFINE_GLV or MSM_FORM_PROJECTIVE");
  CK(cudaSetDevice(ctx->device));
  int c = window_bits > 0 ? window_bits : default_window(ctx->curve, form, n ? n : 1);
  // resident bases with window tables: the tables' window size is the default (shared buckets)
  // (the same for the GLV form of the Weierstrass curves)
  if (window_bits <= 0 && ctx->table_c > 0 && n >= ((size_t)1 << 12) && (te || form == MSM_FORM_AFFINE_GLV)) c = ctx->table_c;
  if (c < 1 || c > 24) return fail(ctx, MSM_E_INVALID, "window_bits out of range [1,24]");
  Timer T(ctx);
  int t0 = -1, t1 = -1;
  const void* d_s = scalars;
  ctx->sc_chunks = 0;
  if (!on_device && n) {
    const size_t sb = scalar_bytes(layout), bytes = n * sb;
    RET_IF(ensure(ctx, ctx->raw_scalars, bytes));
    if (n >= ((size_t)1 << 16) && !ctx->bases_pending) {
      // resident bases: the scalars go up in pieces on the copy stream and the per-scalar phases follow piece by
      // piece (run_*); the copy stream first waits for whatever still reads the buffer on the main stream.
      // (One-shot call: the copy stream is busy with the points, the scalars stay on the main stream.)
      CK(cudaEventRecord(ctx->sc_start, ctx->stream));
      CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->sc_start, 0));
      t0 = T.mark(ctx->copy_stream);
      for (int j = 0; j < msm_b200_ctx::SC_CHUNKS; j++) {
        size_t lo, hi;
        scalar_piece(n, msm_b200_ctx::SC_CHUNKS, j, lo, hi);
        if (hi > lo)
          CK(cudaMemcpyAsync((char*)ctx->raw_scalars.p + lo * sb, (const char*)scalars + lo * sb, (hi - lo) * sb,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->sc_ready[j], ctx->copy_stream));
      }
      t1 = T.mark(ctx->copy_stream);
      ctx->sc_chunks = msm_b200_ctx::SC_CHUNKS;
    } else {
      t0 = T.mark();
      CK(cudaMemcpyAsync(ctx->raw_scalars.p, scalars, bytes, cudaMemcpyHostToDevice, ctx->stream));
      t1 = T.mark();
    }
    d_s = ctx->raw_scalars.p;
  }
  if (tm) {
    memset(tm, 0, sizeof *tm);
  }
  int rc;
  ctx->pending.valid = false;
  ctx->pending.fwd0[0] = ctx->pending.fwd0[1] = -1;
  ctx->pending.window_bits = c;
  ctx->pending.n_windows = 0;
  ctx->pending.shared_buckets = 0;
  if (n == 0)
    rc = ops_of(ctx->curve)->zero_partial(ctx);
  else
    rc = ops_of(ctx->curve)->run(ctx, d_s, n, layout, form, c, tm, digits_dump_dev);
  if (ctx->sc_chunks > 0) {  // an error path left the pieces unconsumed: the main stream must still order behind them
    for (int j = 0; j < ctx->sc_chunks; j++) cudaStreamWaitEvent(ctx->stream, ctx->sc_ready[j], 0);
    ctx->sc_chunks = 0;
  }
  RET_IF(rc);
  RET_IF(wait_for_bases(ctx));  // (no-op unless the MSM never touched the bases, e.g. all-zero scalars)
  ctx->pending.h2d[0] = t0;
  ctx->pending.h2d[1] = t1;
  if (tm) {
    if (ctx->pending.valid && t0 >= 0) tm->h2d_ms = T.ms(t0, t1);
    tm->kernel_launches = ctx->launches;
  }
  return 0;
}

static size_t partial_bytes(int curve) {
  return (curve == MSM_CURVE_ED_ON_BLS12_377 ? 4 : 3) * (size_t)field_limbs(curve) * 4;
}

static int combine_impl(msm_b200_ctx* ctx, const void* partials_dev, int count, msm_b200_point* out) {
  if (!ctx || !out || count < 1) return fail(ctx, MSM_E_INVALID, "bad combine arguments");
  CK(cudaSetDevice(ctx->device));
  return ops_of(ctx->curve)->finalize(ctx, partials_dev, count, out);
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void k_test_field(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  Fe<F> x = fe_one<F>(), y = fe_one<F>(), r;
  if (valid) {
    for (int j = 0; j < F::N; j++) {
      x.v[j] = a[i * F::N + j];
      y.v[j] = b[i * F::N + j];
    }
  }
  if (op == 5) {
    // quad-cooperative inverse: the whole warp works on one argument at a time
    const int lane = threadIdx.x & 31;
    r = x;
    for (int src = 0; src < 32; src++) {
      Fe<F> arg;
      for (int j = 0; j < F::N; j++) arg.v[j] = __shfl_sync(0xffffffffu, x.v[j], src);
      Fe<F> inv = fe_inv_quad(arg);
      if (lane == src) r = inv;
    }
  } else if (!valid) {
    return;
  } else if (op == 0) r = fe_mul(x, y);
  else if (op == 1) r = fe_add(x, y);
  else if (op == 2) r = fe_sub(x, y);
  else if (op == 4) r = fe_sqr(x);
  else r = fe_inv(x);
  if (!valid) return;
  for (int j = 0; j < F::N; j++) out[i * F::N + j] = r.v[j];
}

extern "C" {

const char* msm_b200_global_error(void) { return g_err.c_str(); }

const char* msm_b200_last_error(const msm_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

int msm_b200_create(msm_b200_ctx** out, int curve, int device, void* stream) {
  msm_b200_ctx* ctx = nullptr;
  if (!out) return fail(nullptr, MSM_E_INVALID, "null out pointer");
  *out = nullptr;
  if (curve < 0 || curve > 3) return fail(nullptr, MSM_E_INVALID, "unknown curve");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(nullptr, MSM_E_CUDA, "no such CUDA device");
  CK(cudaSetDevice(device));
  ctx = new msm_b200_ctx();
  ctx->device = device;
  ctx->curve = curve;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // development knobs (launch shapes only, never results); out-of-range values are clamped to what the
  // kernels support: group bits 1..5 (a group must fit a warp), at least one pair / element
  if (const char* e = getenv("MSM_B200_FINISH_ADD")) ctx->finish_add_modmuls = std::max(6.0, atof(e));
  if (const char* e = getenv("MSM_B200_FINISH_ROUND")) ctx->finish_round_modmuls = std::max(1.0, atof(e));
  if (const char* e = getenv("MSM_B200_FINISH_ELEMS")) ctx->finish_max_elems = std::max(1, atoi(e));
  if (const char* e = getenv("MSM_B200_ACC_MIN_PAIRS")) ctx->acc_min_pairs = std::min(std::max(1, atoi(e)), (int)ACC_MAX_PAIRS);
  if (const char* e = getenv("MSM_B200_REDUCE_GB0")) ctx->reduce_gb0 = std::min(std::max(1, atoi(e)), 5);
  if (const char* e = getenv("MSM_B200_REDUCE_WARP_GB")) ctx->reduce_warp_gb = std::min(std::max(1, atoi(e)), 5);
  if (const char* e = getenv("MSM_B200_REDUCE_Q0")) ctx->reduce_quad0 = atoi(e) != 0;
  if (const char* e = getenv("MSM_B200_TABLES")) ctx->tables_enabled = atoi(e) != 0;
  if (const char* e = getenv("MSM_B200_BUCKET_SPLIT")) ctx->bucket_split = std::min(std::max(0, atoi(e)), 1 << 20);
  if (const char* e = getenv("MSM_B200_TABLE_WINDOW")) ctx->table_window = std::min(std::max(0, atoi(e)), 24);
  if (const char* e = getenv("MSM_B200_TABLE_MAX_LOG2N")) ctx->table_max_log2n = std::min(std::max(0, atoi(e)), 26);
  if (const char* e = getenv("MSM_B200_REDUCE_WARP_MIN")) ctx->reduce_warp_min = (size_t)std::max(1ll, atoll(e));
  if (stream) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      return fail(nullptr, MSM_E_CUDA, cudaGetErrorString(e));
    }
    ctx->own_stream = true;
  }
  if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->bases_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->totals_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->sc_start, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->sc_ready[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->sc_ready[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->sc_ready[2], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->sc_ready[3], cudaEventDisableTiming) != cudaSuccess) {
    delete ctx;
    return fail(nullptr, MSM_E_CUDA, "stream / event creation failed");
  }
  cudaError_t e1 = cudaMallocHost((void**)&ctx->h_totals, N_TOTALS * 8);
  cudaError_t e2 = cudaMallocHost((void**)&ctx->h_result, 64 * 4);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    delete ctx;
    return fail(nullptr, MSM_E_CUDA, "pinned allocation failed");
  }
  *out = ctx;
  return 0;
}

void msm_b200_destroy(msm_b200_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  drop_borrowed_bases(ctx);
  for (msm_b200_ctx* b : ctx->borrowers) {  // the loans end here: a later run on a borrower fails with MSM_E_STATE
    cudaStreamSynchronize(b->stream);
    b->bases_owner = nullptr;
    b->n_bases = 0;
    b->table_c = b->table_K = 0;
  }
  ctx->borrowers.clear();
  DevBuf* all[] = {&ctx->bases, &ctx->raw_points, &ctx->raw_scalars, &ctx->hs, &ctx->cnt, &ctx->cntk, &ctx->cursor, &ctx->po,
                   &ctx->totals, &ctx->ent, &ctx->pairkey[0], &ctx->pairkey[1], &ctx->elem[0], &ctx->elem[1],
                   &ctx->prefix, &ctx->red[0], &ctx->red[1], &ctx->partial, &ctx->result, &ctx->buckets, &ctx->rp_tables, &ctx->fin, &ctx->others, &ctx->tilesum};
  for (DevBuf* b : all) release(*b);
  for (int i = 0; i < 8; i++) {
    release(ctx->lvl_pre[i]);
    release(ctx->lvl_tot[i]);
  }
  for (auto e : ctx->ev) cudaEventDestroy(e);
  if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
  if (ctx->h_result) cudaFreeHost(ctx->h_result);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->bases_ready) cudaEventDestroy(ctx->bases_ready);
  if (ctx->totals_ready) cudaEventDestroy(ctx->totals_ready);
  if (ctx->sc_start) cudaEventDestroy(ctx->sc_start);
  for (cudaEvent_t e : ctx->sc_ready)
    if (e) cudaEventDestroy(e);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int msm_b200_set_bases(msm_b200_ctx* ctx, const void* points, size_t n, int layout, int on_device) {
  int rc = set_bases_impl(ctx, points, n, layout, on_device);
  if (rc == 0 && ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return rc;
}

int msm_b200_share_bases(msm_b200_ctx* ctx, msm_b200_ctx* owner) {
  if (!ctx || !owner || ctx == owner) return fail(ctx, MSM_E_INVALID, "bad arguments");
  if (ctx->curve != owner->curve || ctx->device != owner->device)
    return fail(ctx, MSM_E_INVALID, "shared bases need the same curve and device");
  if (owner->bases_owner) return fail(ctx, MSM_E_INVALID, "the lender must own its bases");
  CK(cudaSetDevice(owner->device));
  // everything queued on both contexts first: the lender's ingest / tables, the borrower's last MSM
  if (owner->bases_pending) {
    CK(cudaStreamSynchronize(owner->copy_stream));
    owner->bases_pending = false;
  }
  CK(cudaStreamSynchronize(owner->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (ctx->bases_pending) {
    CK(cudaStreamSynchronize(ctx->copy_stream));
    ctx->bases_pending = false;
  }
  release(ctx->bases);  // the borrower's own record sets are not needed any more
  drop_borrowed_bases(ctx);
  owner->borrowers.push_back(ctx);
  ctx->bases_owner = owner;
  ctx->borrowed_gen = owner->bases_gen;
  ctx->n_bases = owner->n_bases;
  ctx->table_c = owner->table_c;
  ctx->table_K = owner->table_K;
  return 0;
}

int msm_b200_set_bases_async(msm_b200_ctx* ctx, const void* points_host, size_t n, int layout) {
  return set_bases_impl(ctx, points_host, n, layout, 0, /*overlapped=*/true);
}

int msm_b200_run_partial(msm_b200_ctx* ctx, const void* scalars, size_t n, int scalar_layout, int on_device, int form,
                         int window_bits, void* partial_dev, msm_b200_timing* timing) {
  if (!partial_dev) return fail(ctx, MSM_E_INVALID, "null partial pointer");
  auto w0 = std::chrono::steady_clock::now();
  if (ctx) {
    ctx->launches = 0;
    ctx->ev_used = 0;
  }
  RET_IF(run_partial_impl(ctx, scalars, n, scalar_layout, on_device, form, window_bits, timing));
  CK(cudaMemcpyAsync(partial_dev, ctx->partial.p, partial_bytes(ctx->curve), cudaMemcpyDeviceToDevice, ctx->stream));
  if (timing) {  // without a timing struct the partial is ready in stream order; see msm_b200_last_timing
    CK(cudaStreamSynchronize(ctx->stream));
    timing->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
  }
  return 0;
}

int msm_b200_last_timing(msm_b200_ctx* ctx, msm_b200_timing* timing) {
  if (!ctx || !timing) return fail(ctx, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  memset(timing, 0, sizeof *timing);
  return resolve_timing(ctx, timing);
}

size_t msm_b200_partial_bytes(const msm_b200_ctx* ctx) { return ctx ? partial_bytes(ctx->curve) : 0; }

int msm_b200_combine(msm_b200_ctx* ctx, const void* partials_dev, int count, msm_b200_point* out) {
  return combine_impl(ctx, partials_dev, count, out);
}

int msm_b200_run(msm_b200_ctx* ctx, const void* scalars, size_t n, int scalar_layout, int on_device, int form,
                 int window_bits, msm_b200_point* out, msm_b200_timing* timing) {
  if (!out) return fail(ctx, MSM_E_INVALID, "null out pointer");
  auto w0 = std::chrono::steady_clock::now();
  if (ctx) {
    ctx->launches = 0;
    ctx->ev_used = 0;
  }
  // the phase timings are resolved after the final synchronisation, so nothing waits in the middle of the call
  RET_IF(run_partial_impl(ctx, scalars, n, scalar_layout, on_device, form, window_bits, nullptr));
  Timer T(ctx);
  int d0 = T.mark();
  RET_IF(combine_impl(ctx, ctx->partial.p, 1, out));
  int d1 = T.mark();
  CK(cudaStreamSynchronize(ctx->stream));
  if (timing) {
    memset(timing, 0, sizeof *timing);
    RET_IF(resolve_timing(ctx, timing));
    timing->d2h_ms = T.ms(d0, d1);
    timing->kernel_launches = ctx->launches;
    timing->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
  }
  return 0;
}

int msm_b200_msm(msm_b200_ctx* ctx, const void* scalars, int scalar_layout, const void* points, int point_layout,
                 size_t n, int form, int window_bits, msm_b200_point* out, msm_b200_timing* timing) {
  if (!out) return fail(ctx, MSM_E_INVALID, "null out pointer");
  if (!ctx) return fail(nullptr, MSM_E_INVALID, "null context");
  auto w0 = std::chrono::steady_clock::now();
  ctx->launches = 0;
  ctx->ev_used = 0;
  Timer T(ctx);
  // upload + ingest run on the copy stream (overlapped with the scalar phases): time them there
  int i0 = T.mark(ctx->copy_stream);
  RET_IF(set_bases_impl(ctx, points, n, point_layout, 0, /*overlapped=*/true));
  int i1 = T.mark(ctx->copy_stream);
  RET_IF(run_partial_impl(ctx, scalars, n, scalar_layout, 0, form, window_bits, nullptr));
  RET_IF(combine_impl(ctx, ctx->partial.p, 1, out));
  if (timing) {
    memset(timing, 0, sizeof *timing);
    RET_IF(resolve_timing(ctx, timing));
    timing->ingest_ms = T.ms(i0, i1);
    timing->kernel_launches = ctx->launches;
    timing->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - w0).count();
  }
  return 0;
}

size_t msm_b200_point_bytes(const msm_b200_ctx* ctx, int layout) { return ctx ? point_bytes(ctx->curve, layout) : 0; }
size_t msm_b200_scalar_bytes(const msm_b200_ctx* ctx, int layout) {
  (void)ctx;
  return scalar_bytes(layout);
}

int msm_b200_dev_alloc(msm_b200_ctx* ctx, void** out_dev, size_t bytes) {
  if (!ctx || !out_dev) return fail(ctx, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(out_dev, bytes ? bytes : 16));
  return 0;
}
int msm_b200_dev_free(msm_b200_ctx* ctx, void* dev) {
  if (!ctx) return fail(ctx, MSM_E_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  CK(cudaFree(dev));
  return 0;
}
int msm_b200_host_alloc_pinned(void** out_host, size_t bytes) {
  msm_b200_ctx* ctx = nullptr;
  if (!out_host) return fail(nullptr, MSM_E_INVALID, "null out pointer");
  CK(cudaMallocHost(out_host, bytes ? bytes : 16));
  return 0;
}
int msm_b200_host_free_pinned(void* host) {
  msm_b200_ctx* ctx = nullptr;
  CK(cudaFreeHost(host));
  return 0;
}
int msm_b200_host_register(void* host, size_t bytes) {
  msm_b200_ctx* ctx = nullptr;
  if (!host || !bytes) return fail(nullptr, MSM_E_INVALID, "bad arguments");
  CK(cudaHostRegister(host, bytes, cudaHostRegisterPortable));
  return 0;
}
int msm_b200_host_unregister(void* host) {
  msm_b200_ctx* ctx = nullptr;
  if (!host) return fail(nullptr, MSM_E_INVALID, "bad arguments");
  CK(cudaHostUnregister(host));
  return 0;
}
int msm_b200_memcpy_d2h(msm_b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  if (!ctx) return fail(ctx, MSM_E_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int msm_b200_memcpy_h2d(msm_b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  if (!ctx) return fail(ctx, MSM_E_INVALID, "null context");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int msm_b200_random_points_at(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed) {
  if (!ctx || !dst_dev) return fail(ctx, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  return ops_of(ctx->curve)->random_points(ctx, dst_dev, first, n, seed);
}
int msm_b200_random_points(msm_b200_ctx* ctx, void* dst_dev, size_t n, uint64_t seed) {
  return msm_b200_random_points_at(ctx, dst_dev, 0, n, seed);
}

int msm_b200_random_scalars_at(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed) {
  if (!ctx || !dst_dev) return fail(ctx, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(ctx->device));
  return ops_of(ctx->curve)->random_scalars(ctx, dst_dev, first, n, seed);
}
int msm_b200_random_scalars(msm_b200_ctx* ctx, void* dst_dev, size_t n, uint64_t seed) {
  return msm_b200_random_scalars_at(ctx, dst_dev, 0, n, seed);
}

// ---- test / measurement hooks ----
int msm_b200_test_field_op(int device, int field, int op, const uint32_t* a_host, const uint32_t* b_host,
                           uint32_t* out_host, size_t n) {
  msm_b200_ctx* ctx = nullptr;
  if (field < 0 || field > 3 || op < 0 || op > 5 || !a_host || !b_host || !out_host)
    return fail(nullptr, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(device));
  int N = (field == 0 || field == 3) ? 12 : 8;
  size_t bytes = n * N * 4;
  uint32_t *a, *b, *o;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMalloc(&o, bytes));
  CK(cudaMemcpy(a, a_host, bytes, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(b, b_host, bytes, cudaMemcpyHostToDevice));
  unsigned grid = cdiv(n, 64);
  if (field == 0) k_test_field<Bls377Fq><<<grid, 64>>>(op, a, b, o, n);
  else if (field == 1) k_test_field<PallasFp><<<grid, 64>>>(op, a, b, o, n);
  else if (field == 2) k_test_field<Bls377Fr><<<grid, 64>>>(op, a, b, o, n);
  else k_test_field<Bls381Fq><<<grid, 64>>>(op, a, b, o, n);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out_host, o, bytes, cudaMemcpyDeviceToHost));
  cudaFree(a);
  cudaFree(b);
  cudaFree(o);
  return 0;
}

int msm_b200_test_digits(msm_b200_ctx* ctx, const void* scalars_host, size_t n, int window_bits, uint32_t* digits_host,
                         int* n_windows) {
  if (!ctx || !scalars_host || !digits_host || n == 0) return fail(ctx, MSM_E_INVALID, "bad arguments");
  if (ctx->curve == MSM_CURVE_ED_ON_BLS12_377) return fail(ctx, MSM_E_INVALID, "GLV digits need a Weierstrass curve");
  CK(cudaSetDevice(ctx->device));
  int c = window_bits > 0 ? window_bits : 13;
  int b = ctx->curve == MSM_CURVE_BLS12_377_G1 ? 126 : 127;  // Scalar.maxBits of the GLV curves
  int K = (b + 1 + c - 1) / c;
  uint32_t* d_dig;
  CK(cudaMalloc(&d_dig, 2 * n * K * 4));
  size_t saved = ctx->n_bases;
  ctx->n_bases = n;  // digits need no bases
  msm_b200_timing tm;
  int rc = run_partial_impl(ctx, scalars_host, n, MSM_LAYOUT_LE_BYTES, 0, MSM_FORM_AFFINE_GLV, c, &tm, d_dig);
  ctx->n_bases = saved;
  if (rc == 0) {
    cudaMemcpy(digits_host, d_dig, 2 * n * K * 4, cudaMemcpyDeviceToHost);
    if (n_windows) *n_windows = K;
  }
  cudaFree(d_dig);
  return rc;
}

int msm_b200_microbench(int device, int which, int iters, double* ops_per_sec, float* ms_out) {
  msm_b200_ctx* ctx = nullptr;
  if (!ops_per_sec || iters < 1) return fail(nullptr, MSM_E_INVALID, "bad arguments");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  uint32_t* d;
  CK(cudaMalloc(&d, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int blocks = prop.multiProcessorCount * 4, threads = 256;
  double ops = 0;
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    CK(cudaEventRecord(e0));
    switch (which) {
      case 0: k_mb_imad<0><<<blocks, threads>>>(d, iters, 12345u); break;
      case 1: k_mb_imad<1><<<blocks, threads>>>(d, iters, 12345u); break;
      case 2: k_mb_imad<2><<<blocks, threads>>>(d, iters, 12345u); break;
      case 5: k_mb_imad<5><<<blocks, threads>>>(d, iters, 12345u); break;
      case 8: k_mb_imad<8><<<blocks, threads>>>(d, iters, 12345u); break;
      case 3: k_mb_modmul<Bls377Fq><<<blocks, threads>>>(d, iters, 12345u); break;
      case 4: k_mb_modmul<PallasFp><<<blocks, threads>>>(d, iters, 12345u); break;
      case 6: k_mb_modsqr<Bls377Fq><<<blocks, threads>>>(d, iters, 12345u); break;
      case 7: k_mb_modsqr<PallasFp><<<blocks, threads>>>(d, iters, 12345u); break;
      case 9: k_mb_dbl_chain<WeierCurve<Bls377Fq, 3>, true><<<1, 32>>>(d, iters, 12345u); break;
      case 10: k_mb_dbl_chain<WeierCurve<Bls377Fq, 3>, false><<<1, 32>>>(d, iters, 12345u); break;
      case 11: k_mb_dbl_chain<WeierCurve<PallasFp, 15>, true><<<1, 32>>>(d, iters, 12345u); break;
      default: cudaFree(d); return fail(nullptr, MSM_E_INVALID, "unknown benchmark");
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
  }
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  if (which >= 9)
    ops = (double)iters;  // doublings of the one chain
  else if (which == 3 || which == 4 || which == 6 || which == 7)
    ops = (double)blocks * threads * iters * 2.0;
  else
    ops = (double)blocks * threads * iters * (double)MB_INNER * MB_CHAINS;
  *ops_per_sec = ops / (ms * 1e-3);
  if (ms_out) *ms_out = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  return 0;
}

}  // extern "C"

#include "multi.cuh"
