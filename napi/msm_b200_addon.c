/* N-API addon over the C ABI of libmsm_b200.so (include/msm_b200.h): the binding a maintainer of
 * mitschabaude/msm-zprize adds so that `Parallel.msm / msmUnsafe / msmProjective` run on the GPU
 * (ts/msm-b200.ts is the TypeScript side; INTEGRATION.md explains the wiring).
 *
 * Only stable `napi_*` C functions are used.  Node is not part of this image, so this file is compile-
 * checked against napi/stub/node_api.h (tests/test_abi.py) and otherwise untested here; the tested binding
 * is the ctypes one (msm_zprize_b200/_lib.py) over the same entry points.
 *
 *   createContext(curveId, device)                                   -> External
 *   setBases(ctx, memoryBytes, byteOffset, n, layout)        (memoryBytes: the Uint8Array over the wasm memory)
 *   run(ctx, memoryBytes, byteOffset, n, layout, form, windowBits)   -> Promise<{x, y, isZero, timing}>
 *   destroy(ctx)
 *
 * `run` does its work in napi async work (off the event loop): the reference's msm is async as well
 * (src/msm-batched-affine.ts:74-83) and the main thread must stay responsive (src/threads/threads.ts:221-260).
 * Errors: a non-zero return code becomes a thrown Error / rejected Promise carrying msm_b200_last_error()
 * (the reference throws from `assert`, src/util.ts:256, or traps).
 */
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "msm_b200.h"

#define NAPI_OK(call)                                     \
  do {                                                    \
    if ((call) != napi_ok) {                              \
      napi_throw_error(env, "MSM_B200", "napi: " #call);  \
      return NULL;                                        \
    }                                                     \
  } while (0)

static void finalize_ctx(napi_env env, void* data, void* hint) {
  (void)env;
  (void)hint;
  if (data) msm_b200_destroy((msm_b200_ctx*)data);
}

/* a slot that can be emptied by destroy() while the External is still referenced from JS */
typedef struct {
  msm_b200_ctx* ctx;
} ctx_box;

static void finalize_box(napi_env env, void* data, void* hint) {
  ctx_box* box = (ctx_box*)data;
  if (box) {
    finalize_ctx(env, box->ctx, hint);
    free(box);
  }
}

static msm_b200_ctx* unbox(napi_env env, napi_value v) {
  ctx_box* box = NULL;
  if (napi_get_value_external(env, v, (void**)&box) != napi_ok || !box || !box->ctx) {
    napi_throw_error(env, "MSM_B200", "invalid or destroyed context");
    return NULL;
  }
  return box->ctx;
}

static napi_value CreateContext(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  int32_t curve = 0, device = 0;
  NAPI_OK(napi_get_value_int32(env, argv[0], &curve));
  if (argc > 1) NAPI_OK(napi_get_value_int32(env, argv[1], &device));
  ctx_box* box = (ctx_box*)calloc(1, sizeof *box);
  if (!box) {
    napi_throw_error(env, "MSM_B200", "out of memory");
    return NULL;
  }
  if (msm_b200_create(&box->ctx, curve, device, NULL) != MSM_OK) {
    free(box);
    napi_throw_error(env, "MSM_B200", msm_b200_global_error());
    return NULL;
  }
  napi_value ext;
  NAPI_OK(napi_create_external(env, box, finalize_box, NULL, &ext));
  return ext;
}

static napi_value Destroy(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  ctx_box* box = NULL;
  NAPI_OK(napi_get_value_external(env, argv[0], (void**)&box));
  if (box && box->ctx) {
    msm_b200_destroy(box->ctx);
    box->ctx = NULL;
  }
  return NULL;
}

/* (typed array over the wasm memory, byte offset, element count) -> host pointer, bounds checked */
static int region(napi_env env, napi_value view, napi_value off_v, size_t need, uint8_t** out) {
  napi_typedarray_type type;
  size_t len = 0, view_off = 0;
  void* data = NULL;
  napi_value ab;
  if (napi_get_typedarray_info(env, view, &type, &len, &data, &ab, &view_off) != napi_ok || type != napi_uint8_array) {
    napi_throw_type_error(env, "MSM_B200", "expected the Uint8Array view of the wasm memory (memoryBytes)");
    return -1;
  }
  int64_t off = 0;
  if (napi_get_value_int64(env, off_v, &off) != napi_ok || off < 0 || (uint64_t)off > len || need > len - (size_t)off) {
    napi_throw_range_error(env, "MSM_B200", "region outside the memory");
    return -1;
  }
  *out = (uint8_t*)data + off;
  return 0;
}

static napi_value SetBases(napi_env env, napi_callback_info info) {
  size_t argc = 5;
  napi_value argv[5];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  msm_b200_ctx* ctx = unbox(env, argv[0]);
  if (!ctx) return NULL;
  int64_t n = 0;
  int32_t layout = 0;
  NAPI_OK(napi_get_value_int64(env, argv[3], &n));
  NAPI_OK(napi_get_value_int32(env, argv[4], &layout));
  if (n < 0) {
    napi_throw_range_error(env, "MSM_B200", "negative count");
    return NULL;
  }
  uint8_t* p = NULL;
  if (region(env, argv[1], argv[2], (size_t)n * msm_b200_point_bytes(ctx, layout), &p)) return NULL;
  /* synchronous: one H2D copy + the ingest kernel; the caller reuses the bases over many run() calls */
  if (msm_b200_set_bases(ctx, p, (size_t)n, layout, 0) != MSM_OK) napi_throw_error(env, "MSM_B200", msm_b200_last_error(ctx));
  return NULL;
}

typedef struct {
  napi_async_work work;
  napi_deferred deferred;
  napi_ref keepalive; /* the memory view must outlive the copy */
  msm_b200_ctx* ctx;
  const uint8_t* scalars;
  size_t n;
  int layout, form, window_bits;
  int rc;
  msm_b200_point out;
  msm_b200_timing tm;
  char err[256];
} run_job;

static void run_execute(napi_env env, void* data) {
  (void)env;
  run_job* j = (run_job*)data;
  j->rc = msm_b200_run(j->ctx, j->scalars, j->n, j->layout, 0, j->form, j->window_bits, &j->out, &j->tm);
  if (j->rc != MSM_OK) {
    strncpy(j->err, msm_b200_last_error(j->ctx), sizeof j->err - 1);
    j->err[sizeof j->err - 1] = 0;
  }
}

static napi_value bytes_value(napi_env env, const uint8_t* src, size_t n) {
  void* dst = NULL;
  napi_value ab, arr;
  if (napi_create_arraybuffer(env, n, &dst, &ab) != napi_ok) return NULL;
  memcpy(dst, src, n);
  if (napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &arr) != napi_ok) return NULL;
  return arr;
}

static void set_number(napi_env env, napi_value obj, const char* key, double v) {
  napi_value n;
  if (napi_create_double(env, v, &n) == napi_ok) napi_set_named_property(env, obj, key, n);
}

static void run_complete(napi_env env, napi_status status, void* data) {
  run_job* j = (run_job*)data;
  if (status != napi_ok && j->rc == MSM_OK) {
    j->rc = MSM_E_STATE;
    strcpy(j->err, "async work cancelled");
  }
  if (j->rc != MSM_OK) {
    napi_value msg, err;
    napi_create_string_utf8(env, j->err, NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, NULL, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  } else {
    const size_t fb = msm_b200_point_bytes(j->ctx, MSM_LAYOUT_LE_BYTES) / 2; /* 48 or 32 */
    napi_value res, tm, zero;
    napi_create_object(env, &res);
    napi_set_named_property(env, res, "x", bytes_value(env, j->out.x, fb));
    napi_set_named_property(env, res, "y", bytes_value(env, j->out.y, fb));
    napi_get_boolean(env, j->out.is_zero != 0, &zero);
    napi_set_named_property(env, res, "isZero", zero);
    napi_create_object(env, &tm);
    set_number(env, tm, "total", j->tm.total_ms);
    set_number(env, tm, "h2d", j->tm.h2d_ms);
    set_number(env, tm, "prepare points & scalars", j->tm.ingest_ms);
    set_number(env, tm, "slice scalars & count buckets", j->tm.digits_ms);
    set_number(env, tm, "sort points", j->tm.sort_ms);
    set_number(env, tm, "bucket accumulation", j->tm.accumulate_ms);
    set_number(env, tm, "bucket reduction", j->tm.reduce_ms);
    set_number(env, tm, "d2h", j->tm.d2h_ms);
    set_number(env, tm, "windowBits", j->tm.window_bits);
    set_number(env, tm, "windows", j->tm.n_windows);
    set_number(env, tm, "rounds", j->tm.rounds);
    napi_set_named_property(env, res, "timing", tm);
    napi_resolve_deferred(env, j->deferred, res);
  }
  napi_delete_reference(env, j->keepalive);
  napi_delete_async_work(env, j->work);
  free(j);
}

static napi_value Run(napi_env env, napi_callback_info info) {
  size_t argc = 7;
  napi_value argv[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  msm_b200_ctx* ctx = unbox(env, argv[0]);
  if (!ctx) return NULL;
  int64_t n = 0;
  int32_t layout = 0, form = 0, c = 0;
  NAPI_OK(napi_get_value_int64(env, argv[3], &n));
  NAPI_OK(napi_get_value_int32(env, argv[4], &layout));
  NAPI_OK(napi_get_value_int32(env, argv[5], &form));
  if (argc > 6) NAPI_OK(napi_get_value_int32(env, argv[6], &c));
  if (n < 0) {
    napi_throw_range_error(env, "MSM_B200", "negative count");
    return NULL;
  }
  uint8_t* p = NULL;
  if (region(env, argv[1], argv[2], (size_t)n * msm_b200_scalar_bytes(ctx, layout), &p)) return NULL;
  run_job* j = (run_job*)calloc(1, sizeof *j);
  if (!j) {
    napi_throw_error(env, "MSM_B200", "out of memory");
    return NULL;
  }
  j->ctx = ctx;
  j->scalars = p;
  j->n = (size_t)n;
  j->layout = layout;
  j->form = form;
  j->window_bits = c;
  napi_value promise, name;
  if (napi_create_promise(env, &j->deferred, &promise) != napi_ok ||
      napi_create_reference(env, argv[1], 1, &j->keepalive) != napi_ok ||
      napi_create_string_utf8(env, "msm_b200_run", NAPI_AUTO_LENGTH, &name) != napi_ok ||
      napi_create_async_work(env, NULL, name, run_execute, run_complete, j, &j->work) != napi_ok ||
      napi_queue_async_work(env, j->work) != napi_ok) {
    free(j);
    napi_throw_error(env, "MSM_B200", "could not queue the MSM");
    return NULL;
  }
  return promise;
}

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"createContext", NULL, CreateContext, NULL, NULL, NULL, napi_default, NULL},
      {"setBases", NULL, SetBases, NULL, NULL, NULL, napi_default, NULL},
      {"run", NULL, Run, NULL, NULL, NULL, napi_default, NULL},
      {"destroy", NULL, Destroy, NULL, NULL, NULL, napi_default, NULL},
  };
  NAPI_OK(napi_define_properties(env, exports, sizeof props / sizeof props[0], props));
  return exports;
}

NAPI_MODULE(msm_b200, Init)
