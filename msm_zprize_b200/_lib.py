"""ctypes binding of libmsm_b200.so (the C ABI declared in include/msm_b200.h).

This is the tested stand-in for the N-API addon the reference's TypeScript host would load
(INTEGRATION.md): same entry points, same argument meaning.  There is no CPU fallback: if the
CUDA library is missing or no GPU is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# (MSM_B200_LIB: development override, e.g. a build with other compile-time constants under build/)
LIB_PATH = os.environ.get("MSM_B200_LIB") or os.path.join(_HERE, "csrc", "libmsm_b200.so")

# enums (include/msm_b200.h)
CURVE_BLS12_377_G1, CURVE_PALLAS, CURVE_ED_ON_BLS12_377, CURVE_BLS12_381_G1 = 0, 1, 2, 3
FORM_AFFINE_GLV, FORM_PROJECTIVE, FORM_TE_EXTENDED = 0, 1, 2
LAYOUT_LIMB29_MONT, LAYOUT_LE_BYTES = 0, 1
E_INVALID, E_CUDA, E_NOMEM, E_STATE = -1, -2, -3, -4


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("ingest_ms", C.c_float), ("digits_ms", C.c_float),
        ("sort_ms", C.c_float), ("accumulate_ms", C.c_float), ("reduce_ms", C.c_float),
        ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("hot_kernel_ms", C.c_float),
        ("hot_kernel_launches", C.c_int), ("kernel_launches", C.c_int), ("window_bits", C.c_int),
        ("n_windows", C.c_int), ("rounds", C.c_int), ("n_adds", C.c_ulonglong),
        ("shared_buckets", C.c_int), ("fwd_round0_ms", C.c_float), ("fwd_round0_pairs", C.c_uint), ("reserved", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Point(C.Structure):
    _fields_ = [("x", C.c_uint8 * 48), ("y", C.c_uint8 * 48), ("is_zero", C.c_int32)]


class MsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"msm_b200 error {code}: {msg}")
        self.code = code


_lib = None

EXPORTS = [
    "msm_b200_create", "msm_b200_destroy", "msm_b200_last_error", "msm_b200_global_error",
    "msm_b200_set_bases", "msm_b200_share_bases", "msm_b200_set_bases_async", "msm_b200_run", "msm_b200_msm", "msm_b200_run_partial", "msm_b200_last_timing",
    "msm_b200_partial_bytes", "msm_b200_combine", "msm_b200_random_points",
    "msm_b200_random_scalars", "msm_b200_point_bytes", "msm_b200_scalar_bytes",
    "msm_b200_random_points_at", "msm_b200_random_scalars_at",
    "msm_b200_dev_alloc", "msm_b200_dev_free", "msm_b200_host_alloc_pinned",
    "msm_b200_host_free_pinned", "msm_b200_host_register", "msm_b200_host_unregister", "msm_b200_memcpy_d2h", "msm_b200_memcpy_h2d",
    "msm_b200_multi_create", "msm_b200_multi_destroy", "msm_b200_multi_last_error", "msm_b200_multi_devices",
    "msm_b200_multi_gather_kind", "msm_b200_multi_shard_range", "msm_b200_multi_ctx", "msm_b200_multi_set_bases", "msm_b200_multi_share_bases", "msm_b200_multi_set_bases_sharded",
    "msm_b200_multi_run", "msm_b200_multi_run_sharded", "msm_b200_multi_msm", "msm_b200_multi_last_timings",
    "msm_b200_pipeline_create", "msm_b200_pipeline_destroy", "msm_b200_pipeline_last_error", "msm_b200_pipeline_depth",
    "msm_b200_pipeline_set_bases", "msm_b200_pipeline_submit", "msm_b200_pipeline_wait",
]
# include/msm_b200_test.h
TEST_EXPORTS = ["msm_b200_test_field_op", "msm_b200_test_digits", "msm_b200_microbench"]


def lib() -> C.CDLL:
    """Loads the CUDA library; raises loudly if it has not been built (no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MsmError(E_CUDA, f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int
    L.msm_b200_create.argtypes = [C.POINTER(vp), ci, ci, vp]
    L.msm_b200_destroy.argtypes = [vp]
    L.msm_b200_destroy.restype = None
    L.msm_b200_last_error.argtypes = [vp]
    L.msm_b200_last_error.restype = C.c_char_p
    L.msm_b200_global_error.restype = C.c_char_p
    L.msm_b200_set_bases.argtypes = [vp, vp, sz, ci, ci]
    L.msm_b200_set_bases_async.argtypes = [vp, vp, sz, ci]
    L.msm_b200_share_bases.argtypes = [vp, vp]
    L.msm_b200_run.argtypes = [vp, vp, sz, ci, ci, ci, ci, C.POINTER(Point), C.POINTER(Timing)]
    L.msm_b200_msm.argtypes = [vp, vp, ci, vp, ci, sz, ci, ci, C.POINTER(Point), C.POINTER(Timing)]
    L.msm_b200_run_partial.argtypes = [vp, vp, sz, ci, ci, ci, ci, vp, C.POINTER(Timing)]
    L.msm_b200_last_timing.argtypes = [vp, C.POINTER(Timing)]
    L.msm_b200_partial_bytes.argtypes = [vp]
    L.msm_b200_partial_bytes.restype = sz
    L.msm_b200_combine.argtypes = [vp, vp, ci, C.POINTER(Point)]
    L.msm_b200_random_points.argtypes = [vp, vp, sz, C.c_uint64]
    L.msm_b200_random_scalars.argtypes = [vp, vp, sz, C.c_uint64]
    L.msm_b200_random_points_at.argtypes = [vp, vp, sz, sz, C.c_uint64]
    L.msm_b200_random_scalars_at.argtypes = [vp, vp, sz, sz, C.c_uint64]
    L.msm_b200_point_bytes.argtypes = [vp, ci]
    L.msm_b200_point_bytes.restype = sz
    L.msm_b200_scalar_bytes.argtypes = [vp, ci]
    L.msm_b200_scalar_bytes.restype = sz
    L.msm_b200_dev_alloc.argtypes = [vp, C.POINTER(vp), sz]
    L.msm_b200_dev_free.argtypes = [vp, vp]
    L.msm_b200_host_alloc_pinned.argtypes = [C.POINTER(vp), sz]
    L.msm_b200_host_free_pinned.argtypes = [vp]
    L.msm_b200_host_register.argtypes = [vp, sz]
    L.msm_b200_host_unregister.argtypes = [vp]
    L.msm_b200_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.msm_b200_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.msm_b200_multi_create.argtypes = [C.POINTER(vp), ci, C.POINTER(ci), ci]
    L.msm_b200_multi_destroy.argtypes = [vp]
    L.msm_b200_multi_destroy.restype = None
    L.msm_b200_multi_last_error.argtypes = [vp]
    L.msm_b200_multi_last_error.restype = C.c_char_p
    L.msm_b200_multi_devices.argtypes = [vp]
    L.msm_b200_multi_gather_kind.argtypes = [vp]
    L.msm_b200_multi_gather_kind.restype = C.c_char_p
    L.msm_b200_multi_shard_range.argtypes = [sz, ci, ci, C.POINTER(sz), C.POINTER(sz)]
    L.msm_b200_multi_shard_range.restype = None
    L.msm_b200_multi_ctx.argtypes = [vp, ci]
    L.msm_b200_multi_ctx.restype = vp
    L.msm_b200_multi_set_bases.argtypes = [vp, vp, sz, ci]
    L.msm_b200_multi_share_bases.argtypes = [vp, vp]
    L.msm_b200_multi_set_bases_sharded.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), ci]
    L.msm_b200_multi_run.argtypes = [vp, vp, sz, ci, ci, ci, C.POINTER(Point), C.POINTER(Timing)]
    L.msm_b200_multi_run_sharded.argtypes = [vp, C.POINTER(vp), ci, ci, ci, C.POINTER(Point), C.POINTER(Timing)]
    L.msm_b200_multi_msm.argtypes = [vp, vp, ci, vp, ci, sz, ci, ci, C.POINTER(Point), C.POINTER(Timing)]
    L.msm_b200_multi_last_timings.argtypes = [vp, C.POINTER(Timing), ci]
    L.msm_b200_pipeline_create.argtypes = [C.POINTER(vp), ci, C.POINTER(ci), ci, ci]
    L.msm_b200_pipeline_destroy.argtypes = [vp]
    L.msm_b200_pipeline_destroy.restype = None
    L.msm_b200_pipeline_last_error.argtypes = [vp]
    L.msm_b200_pipeline_last_error.restype = C.c_char_p
    L.msm_b200_pipeline_depth.argtypes = [vp]
    L.msm_b200_pipeline_set_bases.argtypes = [vp, vp, sz, ci]
    L.msm_b200_pipeline_submit.argtypes = [vp, vp, sz, ci, ci, ci, C.POINTER(Point), C.POINTER(Timing), C.POINTER(ci)]
    L.msm_b200_pipeline_wait.argtypes = [vp, ci]
    L.msm_b200_test_field_op.argtypes = [ci, ci, ci, vp, vp, vp, sz]
    L.msm_b200_test_digits.argtypes = [vp, vp, sz, ci, vp, C.POINTER(ci)]
    L.msm_b200_microbench.argtypes = [ci, ci, ci, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    _lib = L
    return L


def check(rc: int, ctx=None, multi=None):
    if rc == 0:
        return
    L = lib()
    if multi:
        msg = L.msm_b200_multi_last_error(multi) or b""
    else:
        msg = (L.msm_b200_last_error(ctx) if ctx else L.msm_b200_global_error()) or b""
    raise MsmError(rc, msg.decode("utf-8", "replace"))
