/* N-API addon over the C ABI of libmsm_b200.so (include/msm_b200.h): the binding a maintainer of
 * mitschabaude/msm-zprize adds so that `Parallel.msm / msmUnsafe / msmProjective` run on the GPU(s)
 * (ts/msm-b200.ts is the TypeScript side; INTEGRATION.md explains the wiring).
 *
 * Only stable `napi_*` C functions are used.  Node is not part of this image, so this file is compile-
 * checked against napi/stub/node_api.h (tests/test_abi.py) and otherwise untested here; the tested binding
 * is the ctypes one (msm_zprize_b200/_lib.py) over the same entry points.
 *
 *   createContext(curveId, devices)                                  -> External
 *       devices: a device number or an array of them; the context is always a msm_b200_multi (one device is
 *       the degenerate case), so one `run` fans out over all of them inside the library
 *   setBases(ctx, memoryBytes, byteOffset, n, layout)                -> Promise<void>
 *   run(ctx, memoryBytes, byteOffset, n, layout, form, windowBits)   -> Promise<{x, y, isZero, timing}>
 *       (memoryBytes: the Uint8Array over the wasm memory)
 *   shareBases(ctx, ownerCtx)   run over the bases resident in another context (same curve and devices): several
 *       contexts can then have MSMs in flight over one point set (throughput of a stream of MSMs)
 *   destroy(ctx)
 *   pinMemory(memoryBytes) / unpinMemory(memoryBytes)
 *       page-lock the buffer behind a wasm memory (once, after the memory has its final size), so that uploads from
 *       it are asynchronous DMA at PCIe speed; unpin before the memory grows or is dropped
 *
 * `setBases` and `run` do their work in napi async work (off the event loop): the reference's msm is async as
 * well (src/msm-batched-affine.ts:74-83) and the main thread must stay responsive
 * (src/threads/threads.ts:221-260).  A context is single-caller like one thread pool of the reference: while
 * a job is in flight a second setBases / run / destroy on the same context throws ("context busy") instead of
 * racing on the workspace; every job holds a reference on the context External, so garbage collection cannot
 * finalise a context under a running job.
 * Errors: a non-zero return code becomes a thrown Error / rejected Promise carrying the library's message
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

#define MAX_DEVICES 64

/* a slot that can be emptied by destroy() while the External is still referenced from JS; `busy` is only
 * touched on the JS main thread (set when a job is queued, cleared in its completion callback) */
typedef struct {
  msm_b200_multi* m;
  int busy;
} ctx_box;

static void finalize_box(napi_env env, void* data, void* hint) {
  (void)env;
  (void)hint;
  ctx_box* box = (ctx_box*)data;
  if (box) {
    if (box->m) msm_b200_multi_destroy(box->m);
    free(box);
  }
}

/* the context, or NULL with an exception pending; `for_job`: the caller is about to start work on it */
static ctx_box* unbox(napi_env env, napi_value v, int for_job) {
  ctx_box* box = NULL;
  if (napi_get_value_external(env, v, (void**)&box) != napi_ok || !box || !box->m) {
    napi_throw_error(env, "MSM_B200", "invalid or destroyed context");
    return NULL;
  }
  if (for_job && box->busy) {
    napi_throw_error(env, "MSM_B200", "context busy: await the previous setBases() / run() first");
    return NULL;
  }
  return box;
}

static napi_value CreateContext(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  int32_t curve = 0;
  int devices[MAX_DEVICES] = {0};
  uint32_t n_dev = 1;
  NAPI_OK(napi_get_value_int32(env, argv[0], &curve));
  if (argc > 1) {
    bool is_array = false;
    NAPI_OK(napi_is_array(env, argv[1], &is_array));
    if (is_array) {
      NAPI_OK(napi_get_array_length(env, argv[1], &n_dev));
      if (n_dev < 1 || n_dev > MAX_DEVICES) {
        napi_throw_range_error(env, "MSM_B200", "devices: 1..64 entries");
        return NULL;
      }
      for (uint32_t i = 0; i < n_dev; i++) {
        napi_value e;
        int32_t d = 0;
        NAPI_OK(napi_get_element(env, argv[1], i, &e));
        NAPI_OK(napi_get_value_int32(env, e, &d));
        devices[i] = d;
      }
    } else {
      int32_t d = 0;
      NAPI_OK(napi_get_value_int32(env, argv[1], &d));
      devices[0] = d;
    }
  }
  ctx_box* box = (ctx_box*)calloc(1, sizeof *box);
  if (!box) {
    napi_throw_error(env, "MSM_B200", "out of memory");
    return NULL;
  }
  if (msm_b200_multi_create(&box->m, curve, devices, (int)n_dev) != MSM_OK) {
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
  if (box && box->m) {
    if (box->busy) {
      napi_throw_error(env, "MSM_B200", "context busy: await the running call before destroy()");
      return NULL;
    }
    msm_b200_multi_destroy(box->m);
    box->m = NULL;
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

/* one asynchronous call on a context: setBases (is_run = 0) or run */
typedef struct {
  napi_async_work work;
  napi_deferred deferred;
  napi_ref keep_memory;  /* the memory view must outlive the copy */
  napi_ref keep_context; /* the External must outlive the job (no finalisation under a running MSM) */
  ctx_box* box;
  int is_run;
  const uint8_t* data;
  size_t n;
  int layout, form, window_bits;
  int rc;
  msm_b200_point out;
  msm_b200_timing tm;
  char err[256];
} job_t;

static void job_execute(napi_env env, void* data) {
  (void)env;
  job_t* j = (job_t*)data;
  msm_b200_multi* m = j->box->m;
  if (j->is_run)
    j->rc = msm_b200_multi_run(m, j->data, j->n, j->layout, j->form, j->window_bits, &j->out, &j->tm);
  else
    j->rc = msm_b200_multi_set_bases(m, j->data, j->n, j->layout);
  if (j->rc != MSM_OK) {
    strncpy(j->err, msm_b200_multi_last_error(m), sizeof j->err - 1);
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

static void job_complete(napi_env env, napi_status status, void* data) {
  job_t* j = (job_t*)data;
  j->box->busy = 0;
  if (status != napi_ok && j->rc == MSM_OK) {
    j->rc = MSM_E_STATE;
    strcpy(j->err, "async work cancelled");
  }
  if (j->rc != MSM_OK) {
    napi_value msg, err;
    napi_create_string_utf8(env, j->err, NAPI_AUTO_LENGTH, &msg);
    napi_create_error(env, NULL, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  } else if (!j->is_run) {
    napi_value undef;
    napi_get_undefined(env, &undef);
    napi_resolve_deferred(env, j->deferred, undef);
  } else {
    const size_t fb = msm_b200_point_bytes(msm_b200_multi_ctx(j->box->m, 0), MSM_LAYOUT_LE_BYTES) / 2; /* 48 or 32 */
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
    set_number(env, tm, "devices", msm_b200_multi_devices(j->box->m));
    napi_set_named_property(env, res, "timing", tm);
    napi_resolve_deferred(env, j->deferred, res);
  }
  napi_delete_reference(env, j->keep_memory);
  napi_delete_reference(env, j->keep_context);
  napi_delete_async_work(env, j->work);
  free(j);
}

/* shared by setBases and run: argv = ctx, memoryBytes, byteOffset, n, layout [, form, windowBits] */
static napi_value start_job(napi_env env, napi_callback_info info, int is_run) {
  size_t argc = 7;
  napi_value argv[7];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 5) {
    napi_throw_type_error(env, "MSM_B200", "expected (ctx, memoryBytes, byteOffset, n, layout, ...)");
    return NULL;
  }
  ctx_box* box = unbox(env, argv[0], 1);
  if (!box) return NULL;
  int64_t n = 0;
  int32_t layout = 0, form = 0, c = 0;
  NAPI_OK(napi_get_value_int64(env, argv[3], &n));
  NAPI_OK(napi_get_value_int32(env, argv[4], &layout));
  if (is_run && argc > 5) NAPI_OK(napi_get_value_int32(env, argv[5], &form));
  if (is_run && argc > 6) NAPI_OK(napi_get_value_int32(env, argv[6], &c));
  if (n < 0) {
    napi_throw_range_error(env, "MSM_B200", "negative count");
    return NULL;
  }
  msm_b200_ctx* c0 = msm_b200_multi_ctx(box->m, 0);
  const size_t item = is_run ? msm_b200_scalar_bytes(c0, layout) : msm_b200_point_bytes(c0, layout);
  uint8_t* p = NULL;
  if (region(env, argv[1], argv[2], (size_t)n * item, &p)) return NULL;
  job_t* j = (job_t*)calloc(1, sizeof *j);
  if (!j) {
    napi_throw_error(env, "MSM_B200", "out of memory");
    return NULL;
  }
  j->box = box;
  j->is_run = is_run;
  j->data = p;
  j->n = (size_t)n;
  j->layout = layout;
  j->form = form;
  j->window_bits = c;
  napi_value promise, name;
  if (napi_create_promise(env, &j->deferred, &promise) != napi_ok ||
      napi_create_reference(env, argv[1], 1, &j->keep_memory) != napi_ok) {
    free(j);
    napi_throw_error(env, "MSM_B200", "could not queue the call");
    return NULL;
  }
  if (napi_create_reference(env, argv[0], 1, &j->keep_context) != napi_ok) {
    napi_delete_reference(env, j->keep_memory);
    free(j);
    napi_throw_error(env, "MSM_B200", "could not queue the call");
    return NULL;
  }
  if (napi_create_string_utf8(env, is_run ? "msm_b200_run" : "msm_b200_set_bases", NAPI_AUTO_LENGTH, &name) != napi_ok ||
      napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work) != napi_ok) {
    napi_delete_reference(env, j->keep_memory);
    napi_delete_reference(env, j->keep_context);
    free(j);
    napi_throw_error(env, "MSM_B200", "could not queue the call");
    return NULL;
  }
  box->busy = 1;
  if (napi_queue_async_work(env, j->work) != napi_ok) {
    box->busy = 0;
    napi_delete_async_work(env, j->work);
    napi_delete_reference(env, j->keep_memory);
    napi_delete_reference(env, j->keep_context);
    free(j);
    napi_throw_error(env, "MSM_B200", "could not queue the call");
    return NULL;
  }
  return promise;
}

static napi_value SetBases(napi_env env, napi_callback_info info) { return start_job(env, info, 0); }
static napi_value Run(napi_env env, napi_callback_info info) { return start_job(env, info, 1); }

static napi_value ShareBases(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2) {
    napi_throw_type_error(env, "MSM_B200", "expected (ctx, ownerCtx)");
    return NULL;
  }
  ctx_box* box = unbox(env, argv[0], 1);
  if (!box) return NULL;
  ctx_box* owner = unbox(env, argv[1], 1);
  if (!owner) return NULL;
  if (msm_b200_multi_share_bases(box->m, owner->m) != MSM_OK) napi_throw_error(env, "MSM_B200", msm_b200_multi_last_error(box->m));
  return NULL;
}

static napi_value pin_or_unpin(napi_env env, napi_callback_info info, int pin) {
  size_t argc = 1;
  napi_value argv[1];
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  napi_typedarray_type type;
  size_t len = 0, off = 0;
  void* data = NULL;
  napi_value ab;
  if (argc < 1 || napi_get_typedarray_info(env, argv[0], &type, &len, &data, &ab, &off) != napi_ok || !data || !len) {
    napi_throw_type_error(env, "MSM_B200", "expected the Uint8Array view of the wasm memory (memoryBytes)");
    return NULL;
  }
  int rc = pin ? msm_b200_host_register(data, len) : msm_b200_host_unregister(data);
  if (rc != MSM_OK) napi_throw_error(env, "MSM_B200", msm_b200_global_error());
  return NULL;
}
static napi_value PinMemory(napi_env env, napi_callback_info info) { return pin_or_unpin(env, info, 1); }
static napi_value UnpinMemory(napi_env env, napi_callback_info info) { return pin_or_unpin(env, info, 0); }

static napi_value Init(napi_env env, napi_value exports) {
  napi_property_descriptor props[] = {
      {"createContext", NULL, CreateContext, NULL, NULL, NULL, napi_default, NULL},
      {"setBases", NULL, SetBases, NULL, NULL, NULL, napi_default, NULL},
      {"run", NULL, Run, NULL, NULL, NULL, napi_default, NULL},
      {"destroy", NULL, Destroy, NULL, NULL, NULL, napi_default, NULL},
      {"shareBases", NULL, ShareBases, NULL, NULL, NULL, napi_default, NULL},
      {"pinMemory", NULL, PinMemory, NULL, NULL, NULL, napi_default, NULL},
      {"unpinMemory", NULL, UnpinMemory, NULL, NULL, NULL, napi_default, NULL},
  };
  NAPI_OK(napi_define_properties(env, exports, sizeof props / sizeof props[0], props));
  return exports;
}

NAPI_MODULE(msm_b200, Init)
