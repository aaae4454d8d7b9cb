"""MsmEngine: one GPU + one curve behind the C ABI (include/msm_b200.h).

Host buffers are numpy arrays / bytes in one of the two boundary layouts:
  * LIMB29_MONT -- the reference's in-memory format (src/curve-affine.ts:20-52,
    src/scalar-glv.ts:60-66): what `Parallel.msm(scalarPtr, pointPtr, N)` reads from wasm memory
  * LE_BYTES    -- the compute_msm wire format (src/parallel.ts:97-133,209-249)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _lib as L

CURVES = {"bls12-377": L.CURVE_BLS12_377_G1, "pallas": L.CURVE_PALLAS,
          "ed-on-bls12-377": L.CURVE_ED_ON_BLS12_377, "bls12-381": L.CURVE_BLS12_381_G1}
FIELD_BYTES = {L.CURVE_BLS12_377_G1: 48, L.CURVE_PALLAS: 32, L.CURVE_ED_ON_BLS12_377: 32, L.CURVE_BLS12_381_G1: 48}


@dataclass
class MsmResult:
    x: int
    y: int
    is_zero: bool
    timing: dict


def _as_buffer(a, need: int = 0, what: str = "buffer") -> Tuple[C.c_void_p, object]:
    """Returns (pointer, keepalive) for bytes / bytearray / numpy input without copying numpy.  A buffer
    shorter than `need` bytes is refused here (MSM_E_INVALID) instead of being read past its end by the copy."""
    if isinstance(a, np.ndarray):
        arr = np.ascontiguousarray(a)
    elif isinstance(a, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(a, dtype=np.uint8)
    else:
        raise TypeError(f"unsupported buffer type {type(a)}")
    if arr.nbytes < need:
        raise L.MsmError(L.E_INVALID, f"{what} holds {arr.nbytes} bytes, the call needs {need}")
    return C.c_void_p(arr.ctypes.data), arr


class PinnedBuffer:
    """Page-locked host memory (the analogue of the SharedArrayBuffer the N-API addon reads)."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        L.check(L.lib().msm_b200_host_alloc_pinned(C.byref(self.ptr), nbytes))
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(max(nbytes, 1),))[:nbytes]

    def free(self):
        if self.ptr:
            self.array = None  # the view must not outlive the memory
            L.lib().msm_b200_host_free_pinned(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MsmEngine:
    def __init__(self, curve: str | int, device: int = 0, stream: Optional[int] = None):
        self.curve = CURVES[curve] if isinstance(curve, str) else int(curve)
        self.device = device
        self._ctx = C.c_void_p()
        self._lib = L.lib()
        L.check(self._lib.msm_b200_create(C.byref(self._ctx), self.curve, device, C.c_void_p(stream or 0)))
        self.field_bytes = FIELD_BYTES[self.curve]
        self.default_form = L.FORM_TE_EXTENDED if self.curve == L.CURVE_ED_ON_BLS12_377 else L.FORM_AFFINE_GLV

    # -- lifecycle
    def close(self):
        if self._ctx:
            self._lib.msm_b200_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- sizes
    def point_bytes(self, layout: int) -> int:
        return self._lib.msm_b200_point_bytes(self._ctx, layout)

    def scalar_bytes(self, layout: int) -> int:
        return self._lib.msm_b200_scalar_bytes(self._ctx, layout)

    def partial_bytes(self) -> int:
        return self._lib.msm_b200_partial_bytes(self._ctx)

    # -- inputs
    def set_bases(self, points, n: int, layout: int = L.LAYOUT_LE_BYTES):
        ptr, keep = _as_buffer(points, n * self.point_bytes(layout), "point buffer")
        L.check(self._lib.msm_b200_set_bases(self._ctx, ptr, n, layout, 0), self._ctx)

    def set_bases_async(self, points, n: int, layout: int = L.LAYOUT_LE_BYTES):
        """Queues upload + ingest on the copy stream; the next run()/run_partial() waits for it where it first
        reads a base point.  `points` must stay alive and unchanged until that run has returned."""
        ptr, keep = _as_buffer(points, n * self.point_bytes(layout), "point buffer")
        self._pending_points = keep
        L.check(self._lib.msm_b200_set_bases_async(self._ctx, ptr, n, layout), self._ctx)

    def share_bases(self, owner: "MsmEngine"):
        """Run over the bases resident in `owner` (same device and curve) without copying them: several engines, each
        driven by its own host thread, can then work on independent scalar vectors at the same time."""
        L.check(self._lib.msm_b200_share_bases(self._ctx, owner._ctx), self._ctx)
        self._bases_owner = owner  # keep the lender alive

    def set_bases_device(self, dev_ptr: int, n: int, layout: int = L.LAYOUT_LE_BYTES):
        L.check(self._lib.msm_b200_set_bases(self._ctx, C.c_void_p(dev_ptr), n, layout, 1), self._ctx)

    # -- the MSM
    def _result(self, pt: L.Point, tm: L.Timing) -> MsmResult:
        x = int.from_bytes(bytes(pt.x), "little")
        y = int.from_bytes(bytes(pt.y), "little")
        return MsmResult(x, y, bool(pt.is_zero), tm.as_dict())

    def run(self, scalars, n: int, layout: int = L.LAYOUT_LE_BYTES, form: Optional[int] = None,
            window_bits: int = 0, on_device: bool = False) -> MsmResult:
        pt, tm = L.Point(), L.Timing()
        if on_device:
            ptr, keep = C.c_void_p(int(scalars)), None
        else:
            ptr, keep = _as_buffer(scalars, n * self.scalar_bytes(layout), "scalar buffer")
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_run(self._ctx, ptr, n, layout, int(on_device), form, window_bits,
                                       C.byref(pt), C.byref(tm)), self._ctx)
        return self._result(pt, tm)

    def msm(self, scalars, points, n: int, scalar_layout: int = L.LAYOUT_LE_BYTES,
            point_layout: int = L.LAYOUT_LE_BYTES, form: Optional[int] = None, window_bits: int = 0) -> MsmResult:
        pt, tm = L.Point(), L.Timing()
        sp, k1 = _as_buffer(scalars, n * self.scalar_bytes(scalar_layout), "scalar buffer")
        pp, k2 = _as_buffer(points, n * self.point_bytes(point_layout), "point buffer")
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_msm(self._ctx, sp, scalar_layout, pp, point_layout, n, form, window_bits,
                                       C.byref(pt), C.byref(tm)), self._ctx)
        return self._result(pt, tm)

    def run_partial(self, scalars, n: int, partial_dev_ptr: int, layout: int = L.LAYOUT_LE_BYTES,
                    form: Optional[int] = None, window_bits: int = 0, on_device: bool = False,
                    timing: bool = True) -> Optional[dict]:
        """timing=False: returns without synchronising (the partial is ready in stream order on the engine's
        stream); fetch the phase timings later with last_timing()."""
        tm = L.Timing()
        if on_device:
            ptr, keep = C.c_void_p(int(scalars)), None
        else:
            ptr, keep = _as_buffer(scalars, n * self.scalar_bytes(layout), "scalar buffer")
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_run_partial(self._ctx, ptr, n, layout, int(on_device), form, window_bits,
                                               C.c_void_p(partial_dev_ptr), C.byref(tm) if timing else None),
                self._ctx)
        return tm.as_dict() if timing else None

    def last_timing(self) -> dict:
        tm = L.Timing()
        L.check(self._lib.msm_b200_last_timing(self._ctx, C.byref(tm)), self._ctx)
        return tm.as_dict()

    def combine(self, partials_dev_ptr: int, count: int) -> MsmResult:
        pt = L.Point()
        L.check(self._lib.msm_b200_combine(self._ctx, C.c_void_p(partials_dev_ptr), count, C.byref(pt)), self._ctx)
        return self._result(pt, L.Timing())

    # -- device memory / generators
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        L.check(self._lib.msm_b200_dev_alloc(self._ctx, C.byref(p), nbytes), self._ctx)
        return p.value

    def dev_free(self, ptr: int):
        L.check(self._lib.msm_b200_dev_free(self._ctx, C.c_void_p(ptr)), self._ctx)

    def d2h(self, dev_ptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes, dtype=np.uint8)
        L.check(self._lib.msm_b200_memcpy_d2h(self._ctx, C.c_void_p(out.ctypes.data), C.c_void_p(dev_ptr), nbytes),
                self._ctx)
        return out

    def h2d(self, dev_ptr: int, host) -> None:
        ptr, keep = _as_buffer(host)
        L.check(self._lib.msm_b200_memcpy_h2d(self._ctx, C.c_void_p(dev_ptr), ptr, keep.nbytes), self._ctx)

    def random_points_device(self, dev_ptr: int, n: int, seed: int, first: int = 0):
        """`first`: global index of the first point (a shard of the larger seeded set)."""
        L.check(self._lib.msm_b200_random_points_at(self._ctx, C.c_void_p(dev_ptr), first, n, seed), self._ctx)

    def random_scalars_device(self, dev_ptr: int, n: int, seed: int, first: int = 0):
        L.check(self._lib.msm_b200_random_scalars_at(self._ctx, C.c_void_p(dev_ptr), first, n, seed), self._ctx)

    # -- test hooks
    def test_digits(self, scalars_le: bytes, n: int, window_bits: int) -> np.ndarray:
        kmax = 2 * n * 128
        out = np.zeros(kmax, dtype=np.uint32)
        K = C.c_int()
        ptr, keep = _as_buffer(scalars_le)
        L.check(self._lib.msm_b200_test_digits(self._ctx, ptr, n, window_bits, C.c_void_p(out.ctypes.data),
                                               C.byref(K)), self._ctx)
        return out[: 2 * n * K.value].reshape(2 * n, K.value)


class PipelinedMsm:
    """`depth` engines on one GPU over ONE resident point set (msm_b200_share_bases), each driven by its own host
    thread: independent MSMs submitted back to back overlap -- the inversions and the bucket reduction of one run
    beside the rounds of another.  submit() returns a Future of MsmResult; results come back in submission order
    through map()."""

    def __init__(self, curve: str | int, device: int = 0, depth: int = 2):
        from concurrent.futures import ThreadPoolExecutor
        import queue
        self.engines = [MsmEngine(curve, device=device) for _ in range(depth)]
        self._free: "queue.Queue[MsmEngine]" = queue.Queue()
        for e in self.engines:
            self._free.put(e)
        self._pool = ThreadPoolExecutor(max_workers=depth)

    def set_bases(self, points, n: int, layout: int = L.LAYOUT_LE_BYTES, on_device: bool = False):
        own = self.engines[0]
        if on_device:
            own.set_bases_device(int(points), n, layout)
        else:
            own.set_bases(points, n, layout)
        for e in self.engines[1:]:
            e.share_bases(own)

    def _run(self, scalars, n, kw):
        e = self._free.get()
        try:
            return e.run(scalars, n, **kw)
        finally:
            self._free.put(e)

    def submit(self, scalars, n: int, **kw):
        return self._pool.submit(self._run, scalars, n, kw)

    def map(self, scalar_sets, n: int, **kw):
        return [f.result() for f in [self.submit(s, n, **kw) for s in scalar_sets]]

    def close(self):
        self._pool.shutdown(wait=True)
        for e in reversed(self.engines):  # borrowers first
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class MsmPipeline:
    """msm_b200_pipeline_*: `depth` lanes over one resident point set behind plain calls -- submit() returns a ticket at
    once, wait() returns the MsmResult.  The library owns the lanes' dispatcher threads; the caller needs none."""

    def __init__(self, curve: str | int, devices=(0,), depth: int = 4):
        self.curve = CURVES[curve] if isinstance(curve, str) else int(curve)
        self._lib = L.lib()
        self._p = C.c_void_p()
        devs = list(devices)
        arr = (C.c_int * len(devs))(*devs)
        L.check(self._lib.msm_b200_pipeline_create(C.byref(self._p), self.curve, arr, len(devs), depth))
        self.depth = depth
        self.default_form = L.FORM_TE_EXTENDED if self.curve == L.CURVE_ED_ON_BLS12_377 else L.FORM_AFFINE_GLV
        self._jobs = {}
        fb = FIELD_BYTES[self.curve]
        n29 = 14 if fb == 48 else 9  # 29-bit limbs per field element (src/bigint/field-util.ts:18-42)
        limb29 = 4 * 4 * n29 if self.curve == L.CURVE_ED_ON_BLS12_377 else 2 * 4 * n29 + 4
        self._pbytes = {L.LAYOUT_LE_BYTES: 2 * fb, L.LAYOUT_LIMB29_MONT: limb29}

    def _err(self, rc):
        if rc != 0:
            msg = (self._lib.msm_b200_pipeline_last_error(self._p) or b"").decode("utf-8", "replace")
            raise L.MsmError(rc, msg)

    def set_bases(self, points, n: int, layout: int = L.LAYOUT_LE_BYTES):
        ptr, keep = _as_buffer(points, n * self._pbytes.get(layout, 0), "point buffer")
        self._err(self._lib.msm_b200_pipeline_set_bases(self._p, ptr, n, layout))

    def submit(self, scalars, n: int, layout: int = L.LAYOUT_LE_BYTES, form: Optional[int] = None, window_bits: int = 0) -> int:
        ptr, keep = _as_buffer(scalars, n * (32 if layout == L.LAYOUT_LE_BYTES else 36), "scalar buffer")
        pt, tm, ticket = L.Point(), L.Timing(), C.c_int()
        form = self.default_form if form is None else form
        self._err(self._lib.msm_b200_pipeline_submit(self._p, ptr, n, layout, form, window_bits, C.byref(pt), C.byref(tm),
                                                     C.byref(ticket)))
        self._jobs[ticket.value] = (pt, tm, keep)  # buffers stay alive until wait()
        return ticket.value

    def wait(self, ticket: int) -> MsmResult:
        pt, tm, _ = self._jobs.pop(ticket)
        self._err(self._lib.msm_b200_pipeline_wait(self._p, ticket))
        return MsmResult(int.from_bytes(bytes(pt.x), "little"), int.from_bytes(bytes(pt.y), "little"), bool(pt.is_zero),
                         tm.as_dict())

    def close(self):
        if self._p:
            self._lib.msm_b200_pipeline_destroy(self._p)
            self._p = C.c_void_p()
            self._jobs.clear()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiMsmEngine:
    """Several GPUs behind one call (msm_b200_multi_*): the library shards the points by range, drives every
    device from its own host thread, gathers the partial points (NCCL / peer copies) and adds them on
    device 0.  No torch, no torchrun: this is what the N-API addon calls with `devices[]`."""

    def __init__(self, curve: str | int, devices):
        self.curve = CURVES[curve] if isinstance(curve, str) else int(curve)
        self.devices = list(devices)
        self._lib = L.lib()
        self._m = C.c_void_p()
        arr = (C.c_int * len(self.devices))(*self.devices)
        L.check(self._lib.msm_b200_multi_create(C.byref(self._m), self.curve, arr, len(self.devices)))
        self.field_bytes = FIELD_BYTES[self.curve]
        self.default_form = L.FORM_TE_EXTENDED if self.curve == L.CURVE_ED_ON_BLS12_377 else L.FORM_AFFINE_GLV
        self.gather_kind = (self._lib.msm_b200_multi_gather_kind(self._m) or b"").decode()
        # per-device views (generators, device memory); owned by the multi context
        self.shards = []
        for i in range(len(self.devices)):
            e = MsmEngine.__new__(MsmEngine)
            e.curve, e.device, e._lib = self.curve, self.devices[i], self._lib
            e._ctx = C.c_void_p(self._lib.msm_b200_multi_ctx(self._m, i))
            e.field_bytes, e.default_form = self.field_bytes, self.default_form
            e.close = lambda: None  # not ours to destroy
            self.shards.append(e)

    def close(self):
        if self._m:
            for e in self.shards:
                e._ctx = C.c_void_p()
            self._lib.msm_b200_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _result(self, pt, tm):
        return MsmResult(int.from_bytes(bytes(pt.x), "little"), int.from_bytes(bytes(pt.y), "little"), bool(pt.is_zero),
                         tm.as_dict())

    def set_bases(self, points, n: int, layout: int = L.LAYOUT_LE_BYTES):
        ptr, keep = _as_buffer(points, n * self.shards[0].point_bytes(layout), "point buffer")
        L.check(self._lib.msm_b200_multi_set_bases(self._m, ptr, n, layout), multi=self._m)

    def share_bases(self, owner: "MultiMsmEngine"):
        """Run over the bases resident in `owner` (same curve and device list): several multi engines, each driven from
        its own host thread, can then have MSMs in flight over one point set."""
        L.check(self._lib.msm_b200_multi_share_bases(self._m, owner._m), multi=self._m)
        self._bases_owner = owner

    def set_bases_sharded(self, dev_ptrs, counts, layout: int = L.LAYOUT_LE_BYTES):
        p = (C.c_void_p * len(dev_ptrs))(*[C.c_void_p(int(x)) for x in dev_ptrs])
        c = (C.c_size_t * len(counts))(*counts)
        L.check(self._lib.msm_b200_multi_set_bases_sharded(self._m, p, c, layout), multi=self._m)

    def run(self, scalars, n: int, layout: int = L.LAYOUT_LE_BYTES, form: Optional[int] = None,
            window_bits: int = 0) -> MsmResult:
        pt, tm = L.Point(), L.Timing()
        ptr, keep = _as_buffer(scalars, n * self.shards[0].scalar_bytes(layout), "scalar buffer")
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_multi_run(self._m, ptr, n, layout, form, window_bits, C.byref(pt), C.byref(tm)),
                multi=self._m)
        return self._result(pt, tm)

    def run_sharded(self, dev_ptrs, layout: int = L.LAYOUT_LE_BYTES, form: Optional[int] = None,
                    window_bits: int = 0) -> MsmResult:
        pt, tm = L.Point(), L.Timing()
        p = (C.c_void_p * len(dev_ptrs))(*[C.c_void_p(int(x)) for x in dev_ptrs])
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_multi_run_sharded(self._m, p, layout, form, window_bits, C.byref(pt), C.byref(tm)),
                multi=self._m)
        return self._result(pt, tm)

    def msm(self, scalars, points, n: int, scalar_layout: int = L.LAYOUT_LE_BYTES, point_layout: int = L.LAYOUT_LE_BYTES,
            form: Optional[int] = None, window_bits: int = 0) -> MsmResult:
        pt, tm = L.Point(), L.Timing()
        sp, k1 = _as_buffer(scalars, n * self.shards[0].scalar_bytes(scalar_layout), "scalar buffer")
        pp, k2 = _as_buffer(points, n * self.shards[0].point_bytes(point_layout), "point buffer")
        form = self.default_form if form is None else form
        L.check(self._lib.msm_b200_multi_msm(self._m, sp, scalar_layout, pp, point_layout, n, form, window_bits,
                                             C.byref(pt), C.byref(tm)), multi=self._m)
        return self._result(pt, tm)

    def last_timings(self):
        arr = (L.Timing * len(self.devices))()
        L.check(self._lib.msm_b200_multi_last_timings(self._m, arr, len(self.devices)), multi=self._m)
        return [t.as_dict() for t in arr]


def test_field_op(device: int, field: int, op: int, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty_like(a)
    L.check(L.lib().msm_b200_test_field_op(device, field, op, C.c_void_p(a.ctypes.data), C.c_void_p(b.ctypes.data),
                                           C.c_void_p(out.ctypes.data), a.shape[0]))
    return out


def microbench(device: int, which: int, iters: int = 64) -> Tuple[float, float]:
    ops = C.c_double()
    ms = C.c_float()
    L.check(L.lib().msm_b200_microbench(device, which, iters, C.byref(ops), C.byref(ms)))
    return ops.value, ms.value
