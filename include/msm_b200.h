/* msm_b200.h -- C ABI of libmsm_b200.so, the B200-native MSM engine.
 *
 * Drop-in boundary for the reference's MSM entry points (scalars + affine points in, one curve
 * point out).  Every entry point names the reference interface it replaces (paths relative to
 * the reference repo).  Plain pointers and sizes only; no C++/torch types.  All functions return
 * 0 on success or a negative MSM_E_* code; msm_b200_last_error() gives a message.  Nothing here
 * throws or aborts across the ABI (the reference's JS exceptions / wasm traps become codes).
 *
 * One context = one GPU + one curve.  A context is single-caller (not re-entrant), like one
 * thread-pool instance of the reference (src/threads/threads.ts:132-277).
 */
#ifndef MSM_B200_H
#define MSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msm_b200_ctx msm_b200_ctx;

/* curves: src/concrete/bls12-377.params.ts, pasta.params.ts, ed-on-bls12-377.params.ts */
enum msm_b200_curve {
  MSM_CURVE_BLS12_377_G1 = 0, /* short Weierstrass a=0, GLV, 377-bit base field */
  MSM_CURVE_PALLAS = 1,       /* short Weierstrass a=0, GLV, 255-bit base field */
  MSM_CURVE_ED_ON_BLS12_377 = 2, /* twisted Edwards a=-1, 253-bit base field */
  MSM_CURVE_BLS12_381_G1 = 3     /* short Weierstrass a=0, GLV, 381-bit base field (src/concrete/bls12-381.params.ts) */
};

/* algorithm ("form"): which of the reference's MSMs the call mirrors */
enum msm_b200_form {
  MSM_FORM_AFFINE_GLV = 0, /* createMsm().msm / msmUnsafe, src/msm-batched-affine.ts:74-328,573-587 */
  MSM_FORM_PROJECTIVE = 1, /* Parallel.msmProjective -> msmBasic, src/parallel.ts:69-87, src/msm-basic.ts:45-176 */
  MSM_FORM_TE_EXTENDED = 2 /* TwistedEdwards Parallel.msm -> msmBasic, src/parallel.ts:193-199 */
};

/* data layouts at the boundary */
enum msm_b200_layout {
  /* The reference's in-memory (wasm) format.  Points: src/curve-affine.ts:20-52 (x | y | u8
   * isNonZero + 3 pad; 2*4*n+4 bytes, n = 14 or 9 limbs of 29 bits in u32 words, Montgomery
   * R = 2^(29n), coordinates may be unreduced in [0,2p)); twisted Edwards: X|Y|Z|T, 4*4*n bytes,
   * Z = R mod p (src/curve-twisted-edwards.ts:30-31).  Scalars: 9 x 29-bit limbs in u32 words,
   * plain integers < q (src/scalar-glv.ts:60-66). */
  MSM_LAYOUT_LIMB29_MONT = 0,
  /* The compute_msm wire format: canonical little-endian bytes.  Points x|y, 48|48 bytes
   * (BLS12-377) or 32|32 (Pallas, ed-on-bls12-377): src/parallel.ts:97-116,209-232.
   * Scalars 32 bytes LE: src/parallel.ts:119-133.  No representation of the zero point
   * (src/parallel.ts:107-108). */
  MSM_LAYOUT_LE_BYTES = 1
};

enum msm_b200_error {
  MSM_OK = 0,
  MSM_E_INVALID = -1, /* bad argument (replaces `assert`, src/util.ts:256) */
  MSM_E_CUDA = -2,    /* CUDA runtime failure, incl. no device */
  MSM_E_NOMEM = -3,   /* device allocation failed (replaces the memory-overflow exception,
                         src/wasm/memory-helpers.ts:224-236) */
  MSM_E_STATE = -4    /* call order (e.g. run before set_bases) */
};

/* Per-phase timings in milliseconds (CUDA events), the analogue of the reference's tic/toc log
 * (src/msm-common.ts:192-230) returned by msm(..., verboseTiming=true). */
typedef struct msm_b200_timing {
  float h2d_ms;       /* scalars (and points, for the one-shot call) host -> device */
  float ingest_ms;    /* layout conversion: "prepare points & scalars" */
  float digits_ms;    /* GLV decompose + signed digits + histogram: "slice scalars & count buckets" */
  float sort_ms;      /* offsets + scatter: "integrate bucket counts", "sort points" */
  float accumulate_ms;/* "bucket accumulation" */
  float reduce_ms;    /* "bucket reduction", "partition sum", "final sum" + normalisation */
  float d2h_ms;
  float total_ms;     /* whole call, host clock */
  /* dominant kernel (batched-affine backward pass / bucket accumulate) */
  float hot_kernel_ms;      /* sum of its launch durations in this call */
  int hot_kernel_launches;
  int kernel_launches;      /* all kernels launched by this call */
  int window_bits;          /* c actually used */
  int n_windows;            /* K */
  int rounds;               /* batched-affine tree rounds */
  unsigned long long n_adds;/* point additions finished inside the dominant kernel's launches */
  int shared_buckets;       /* 1: resident window tables 2^(kc) G were used, all windows share one bucket set */
  float fwd_round0_ms;      /* batched-affine path: the round-0 forward pass (random gathers of base points, HBM-bound) */
  unsigned fwd_round0_pairs;/*   and the point pairs it went through */
  int reserved;
} msm_b200_timing;

/* Result point, canonical affine (the normalisation that defines parity: Projective.toAffine +
 * Affine.toBigint, src/curve-projective.ts:335-349, src/curve-affine.ts:220-233; twisted Edwards:
 * src/curve-twisted-edwards.ts:369-386).  x, y little-endian, 48 bytes used for BLS12-377, 32
 * for the others (rest zero).  is_zero: Weierstrass point at infinity (x = y = 0 then); the
 * twisted-Edwards zero is the ordinary point (0, 1) with is_zero = 1 as a convenience. */
typedef struct msm_b200_point {
  uint8_t x[48];
  uint8_t y[48];
  int32_t is_zero;
} msm_b200_point;

/* -- lifecycle ------------------------------------------------------------------------------
 * replaces Weierstrass.create / TwistedEdwards.create + startThreads (src/parallel.ts:40-177,
 * 179-289, 291-315): builds the per-curve engine on CUDA device `device`.
 * `stream`: a cudaStream_t to run on (e.g. torch's current stream), or NULL for an own stream. */
int msm_b200_create(msm_b200_ctx** out, int curve, int device, void* stream);
/* replaces stopThreads (src/parallel.ts:317-320) */
void msm_b200_destroy(msm_b200_ctx* ctx);
const char* msm_b200_last_error(const msm_b200_ctx* ctx);
/* library-level message for failures before a context exists */
const char* msm_b200_global_error(void);

/* -- inputs ---------------------------------------------------------------------------------
 * Uploads and prepares the point set (kept resident, like points living in wasm memory across
 * benchmark iterations, scripts/msm-weierstrass.ts:19,29-33).  Does on the device what
 * preparePointsAndScalars does per point (copy, endomorphism; src/msm-batched-affine.ts:338-409)
 * and, for LE_BYTES, what Parallel.pointsFromBytes does (src/parallel.ts:97-116,209-232).
 * `points` is host memory unless `on_device` != 0.
 * For 2^14 <= n <= 2^25 points (2^13 for twisted Edwards) this call also builds the window tables 2^(kc) G_i of
 * the set (K - 1 more record sets in device memory, about three MSMs of work, once): later msm_b200_run calls
 * with the default window then add the digits of all windows into one shared set of buckets
 * (timing.shared_buckets = 1; same point; K times fewer buckets to reduce, no Horner step, and wider windows --
 * fewer additions -- pay earlier).  MSM_B200_TABLES=0 in the environment turns this off; an explicit window_bits
 * other than the tables' and the one-shot msm_b200_msm use the classic layout (one bucket set per window). */
int msm_b200_set_bases(msm_b200_ctx* ctx, const void* points, size_t n, int layout, int on_device);
/* Lets `ctx` run its MSMs over the bases resident in `owner` (same device, same curve) without a second copy:
 * several contexts -- each with its own stream and workspace, each driven by its own host thread -- can then work
 * on independent scalar vectors over ONE point set at the same time, so that the latency-bound phases of one MSM
 * (inversions, bucket reduction) overlap the throughput-bound rounds of another (bench.py `pipelined`: +32 %
 * MSMs per second at 2^18 points with two contexts).  The reference has one MSM in flight per thread pool.
 * The loan ends when `ctx` gets bases of its own; after `owner` replaces ITS bases, runs on `ctx` fail with
 * MSM_E_STATE until msm_b200_share_bases is called again; destroying `owner` ends the loan the same way (the borrower
 * is left without bases).  Not to be called while either context has an MSM running on another thread. */
int msm_b200_share_bases(msm_b200_ctx* ctx, msm_b200_ctx* owner);
/* Same for host points, without waiting: the copy and the ingest kernel are queued on the context's copy
 * stream and the next run / run_partial waits for them only where it first reads a base point, i.e. behind its
 * own scalar upload, GLV and sort phases (what msm_b200_msm does inside one call).  `points_host` must stay
 * valid and unchanged until that run has returned; pinned memory makes the copy asynchronous. */
int msm_b200_set_bases_async(msm_b200_ctx* ctx, const void* points_host, size_t n, int layout);

/* -- the MSM --------------------------------------------------------------------------------
 * replaces Parallel.msm / Parallel.msmUnsafe / Parallel.msmProjective
 * (src/msm-batched-affine.ts:74-83,573-587; src/parallel.ts:69-87; src/msm-basic.ts:34-43) over
 * the resident bases: sum_i scalars[i] * bases[i], i < n <= n_bases.
 * `window_bits`: the reference's options.c (0 = engine default).  `form`: msm_b200_form.
 * `scalars` is host memory unless `on_device` != 0.  `timing` may be NULL. */
int msm_b200_run(msm_b200_ctx* ctx, const void* scalars, size_t n, int scalar_layout, int on_device,
                 int form, int window_bits, msm_b200_point* out, msm_b200_timing* timing);

/* One-shot call with host buffers for both inputs: set_bases + run, the shape of
 * compute_msm(points, scalars) (scripts/zprize23/submission-bls377.ts:20-65, submission.ts:19-35)
 * and of a cold Parallel.msm call. */
int msm_b200_msm(msm_b200_ctx* ctx, const void* scalars, int scalar_layout, const void* points,
                 int point_layout, size_t n, int form, int window_bits, msm_b200_point* out,
                 msm_b200_timing* timing);

/* Same as msm_b200_run but leaves the result on the device: writes one partial (internal
 * representation, msm_b200_partial_bytes() bytes) to `partial_dev` (device memory, e.g. a torch
 * tensor that is then all-gathered with NCCL).  Multi-GPU: each rank owns a contiguous point
 * range (the GPU analogue of range(), src/threads/threads.ts:354-359). */
int msm_b200_run_partial(msm_b200_ctx* ctx, const void* scalars, size_t n, int scalar_layout,
                         int on_device, int form, int window_bits, void* partial_dev,
                         msm_b200_timing* timing);
size_t msm_b200_partial_bytes(const msm_b200_ctx* ctx);
/* With timing == NULL, msm_b200_run_partial returns WITHOUT synchronising: the partial is ready in stream
 * order on the context's stream, so a collective queued on that stream can follow at once.  The phase
 * timings of that last call (the `log` of createLog, src/msm-common.ts:192-230) can be fetched afterwards
 * with this function, which waits for the stream first. */
int msm_b200_last_timing(msm_b200_ctx* ctx, msm_b200_timing* timing);
/* Adds `count` gathered partials (device memory) and normalises: the "partition sum / final sum"
 * of src/msm-batched-affine.ts:299-322 across GPUs. */
int msm_b200_combine(msm_b200_ctx* ctx, const void* partials_dev, int count, msm_b200_point* out);

/* -- synthetic inputs on the device ----------------------------------------------------------
 * replaces Parallel.randomPointsFast / randomScalars (src/curve-random.ts:24-91,151-194) with
 * seeded generators (the reference has no seeds, src/util.ts:226-233): writes `n` points /
 * scalars in LE_BYTES layout into device memory `dst_dev` (n * point_bytes / n * 32 bytes). */
int msm_b200_random_points(msm_b200_ctx* ctx, void* dst_dev, size_t n, uint64_t seed);
int msm_b200_random_scalars(msm_b200_ctx* ctx, void* dst_dev, size_t n, uint64_t seed);
/* The same generators for a RANGE of a larger seeded set: element j of the output is element `first + j` of
 * the set that msm_b200_random_points / _scalars(seed) define -- each GPU generates its own shard. */
int msm_b200_random_points_at(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed);
int msm_b200_random_scalars_at(msm_b200_ctx* ctx, void* dst_dev, size_t first, size_t n, uint64_t seed);
size_t msm_b200_point_bytes(const msm_b200_ctx* ctx, int layout);
size_t msm_b200_scalar_bytes(const msm_b200_ctx* ctx, int layout);

/* -- device memory helpers for hosts without a CUDA binding (the N-API addon) ---------------- */
int msm_b200_dev_alloc(msm_b200_ctx* ctx, void** out_dev, size_t bytes);
int msm_b200_dev_free(msm_b200_ctx* ctx, void* dev);
int msm_b200_host_alloc_pinned(void** out_host, size_t bytes);
int msm_b200_host_free_pinned(void* host);
/* Page-locks memory the caller already owns (the buffer behind a WebAssembly.Memory): uploads from it then run as
 * asynchronous DMA at PCIe speed instead of being staged through a bounce buffer.  Undo before the memory is freed
 * or grown (a grown wasm memory has a new buffer). */
int msm_b200_host_register(void* host, size_t bytes);
int msm_b200_host_unregister(void* host);
int msm_b200_memcpy_d2h(msm_b200_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int msm_b200_memcpy_h2d(msm_b200_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);

/* -- several GPUs behind ONE call ------------------------------------------------------------
 * The reference's msm is one call that fans out internally: every pool thread takes a static range of the
 * points (range(), src/threads/threads.ts:354-359) and the main thread adds the partition sums
 * (src/msm-batched-affine.ts:294-322).  msm_b200_multi_* is that shape over the GPUs of one box: the point set
 * shards by contiguous range (device g owns [g * per, (g + 1) * per), per = ceil(n / n_dev)), one host thread per
 * device inside the library drives the complete single-GPU pipeline on its range, the n_dev partial points
 * (<= 144 bytes each) are gathered with one ncclAllGather over NVLink (libnccl.so.2 opened at run time; peer
 * copies when it is absent or MSM_B200_GATHER=peer) and device 0 adds them and normalises.  The result is
 * identical to the single-GPU one for every n_dev.  A multi context is single-caller like msm_b200_ctx (calls
 * are serialised by an internal mutex). */
typedef struct msm_b200_multi msm_b200_multi;
int msm_b200_multi_create(msm_b200_multi** out, int curve, const int* devices, int n_dev);
void msm_b200_multi_destroy(msm_b200_multi* m);
const char* msm_b200_multi_last_error(const msm_b200_multi* m);
int msm_b200_multi_devices(const msm_b200_multi* m);
/* "ncclAllGather (NCCL <version>, ncclCommInitAll)" | "peer copies (cudaMemcpyPeerAsync)" | "none (single device)" */
const char* msm_b200_multi_gather_kind(const msm_b200_multi* m);
/* the range of a set of n points that device i of n_dev owns: [first, first + count), ceil(n / n_dev) points each
 * like range() (src/threads/threads.ts:354-359); pure host arithmetic, needs no GPU */
void msm_b200_multi_shard_range(size_t n, int i, int n_dev, size_t* first, size_t* count);
/* the per-device context i (generators, device memory helpers); owned by `m` */
msm_b200_ctx* msm_b200_multi_ctx(msm_b200_multi* m, int i);
/* msm_b200_set_bases over all devices: host points, range-sharded by the library */
int msm_b200_multi_set_bases(msm_b200_multi* m, const void* points_host, size_t n, int layout);
/* msm_b200_share_bases on every device: `m` runs over the bases resident in `owner` (same curve, same device list) */
int msm_b200_multi_share_bases(msm_b200_multi* m, msm_b200_multi* owner);
/* bases already resident on the devices: shard i = n_per_dev[i] points at points_dev[i] in device i's memory
 * (global order: shard 0, shard 1, ...) */
int msm_b200_multi_set_bases_sharded(msm_b200_multi* m, const void* const* points_dev, const size_t* n_per_dev, int layout);
/* msm_b200_run over all devices: host scalars (n <= resident bases; the first n points of the global order) */
int msm_b200_multi_run(msm_b200_multi* m, const void* scalars_host, size_t n, int scalar_layout, int form, int window_bits,
                       msm_b200_point* out, msm_b200_timing* timing);
/* the same with the scalars already on the devices, one shard per device matching the resident bases */
int msm_b200_multi_run_sharded(msm_b200_multi* m, const void* const* scalars_dev, int scalar_layout, int form, int window_bits,
                               msm_b200_point* out, msm_b200_timing* timing);
/* msm_b200_msm over all devices (one-shot: points and scalars from the host every call) */
int msm_b200_multi_msm(msm_b200_multi* m, const void* scalars_host, int scalar_layout, const void* points_host, int point_layout,
                       size_t n, int form, int window_bits, msm_b200_point* out, msm_b200_timing* timing);
/* `timing` of the calls above = the slowest device's phases (launches and additions summed, total_ms = the
 * call's host clock); this returns every device's own phases of the last call */
int msm_b200_multi_last_timings(msm_b200_multi* m, msm_b200_timing* per_device, int count);

/* -- several MSMs in flight behind plain calls -------------------------------------------------
 * `depth` lanes over ONE resident point set (lane 0 owns the bases, the others borrow them, msm_b200_share_bases),
 * each with a dispatcher thread inside the library: submit() hands a scalar vector to the next lane and returns a
 * ticket at once, wait() blocks until that MSM is done and its point is in the caller's buffer.  Scalars, `out` and
 * `timing` must stay valid until wait() has returned for the ticket; at most 4 * depth tickets may be outstanding.
 * Independent MSMs overlap on the GPU (the inversions and the bucket reduction of one beside the rounds of another):
 * 2^18 points, four lanes: 1.55 instead of 2.25 ms per MSM.  Each lane is a multi context, so `devices` may list
 * several GPUs. */
typedef struct msm_b200_pipeline msm_b200_pipeline;
int msm_b200_pipeline_create(msm_b200_pipeline** out, int curve, const int* devices, int n_dev, int depth);
void msm_b200_pipeline_destroy(msm_b200_pipeline* p);
const char* msm_b200_pipeline_last_error(const msm_b200_pipeline* p);
int msm_b200_pipeline_depth(const msm_b200_pipeline* p);
/* waits for the MSMs in flight, then uploads and prepares the points (as msm_b200_multi_set_bases) */
int msm_b200_pipeline_set_bases(msm_b200_pipeline* p, const void* points_host, size_t n, int layout);
int msm_b200_pipeline_submit(msm_b200_pipeline* p, const void* scalars_host, size_t n, int scalar_layout, int form, int window_bits,
                             msm_b200_point* out, msm_b200_timing* timing, int* ticket);
int msm_b200_pipeline_wait(msm_b200_pipeline* p, int ticket);

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_H */
