"""ctypes wrapper of oracle/libmsm_port.so (the multi-threaded C++ port of the reference's MSM).
TEST / BASELINE INFRASTRUCTURE ONLY -- see the header of msm_port.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import bigint_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmsm_port.so")


class PortParams(C.Structure):
    _fields_ = [(k, C.c_uint8 * 64) for k in ("p", "q", "beta", "k2d", "v00", "v01", "v10", "v11", "m0", "m1")] + \
               [(k, C.c_int32) for k in ("v00_neg", "v01_neg", "v10_neg", "v11_neg", "m0_neg", "m1_neg", "glv_m",
                                         "glv_k", "glv_max_bits", "scalar_bits", "is_te", "b3")]


def _lib():
    if not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE])
    L = C.CDLL(LIB_PATH)
    L.port_create.restype = C.c_void_p
    L.port_create.argtypes = [C.POINTER(PortParams), C.c_int]
    L.port_destroy.argtypes = [C.c_void_p]
    L.port_point_words.restype = C.c_size_t
    L.port_point_words.argtypes = [C.c_void_p]
    L.port_points_from_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    L.port_msm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_void_p,
                           C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.port_random_scalars.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int]
    L.port_random_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64,
                                     C.c_int, C.c_int]
    L.port_mul_ns.restype = C.c_double
    L.port_mul_ns.argtypes = [C.c_void_p, C.c_int]
    return L


def _set(field, value: int):
    b = int(value).to_bytes(64, "little")
    for i in range(64):
        field[i] = b[i]


class Port:
    """One curve of the CPU port.  curve: 'bls12-377' | 'pallas' | 'bls12-381' | 'ed-on-bls12-377'."""

    def __init__(self, curve: str):
        self.L = _lib()
        pp = PortParams()
        self.curve = curve
        if curve == "ed-on-bls12-377":
            P = O.ED_ON_BLS12_377
            _set(pp.p, P.p), _set(pp.q, P.q), _set(pp.k2d, 2 * P.d % P.p)
            pp.is_te, pp.scalar_bits = 1, O.log2(P.q)
            self.nbytes, n29 = 32, O.montgomery_params(P.p).n
            self.field_bits = O.log2(P.p)
        else:
            P = {"bls12-377": O.BLS12_377, "pallas": O.PALLAS, "bls12-381": O.BLS12_381}[curve]
            g = O.glv_params(P.q, P.lam)
            _set(pp.p, P.p), _set(pp.q, P.q), _set(pp.beta, P.beta)
            for k in ("v00", "v01", "v10", "v11", "m0", "m1"):
                v = getattr(g, k)
                _set(getattr(pp, k), abs(v))
                setattr(pp, k + "_neg", 1 if v < 0 else 0)
            pp.glv_m, pp.glv_k, pp.glv_max_bits = g.m, g.k, g.max_bits
            pp.scalar_bits, pp.is_te, pp.b3 = O.log2(P.q), 0, 3 * P.b
            self.nbytes, n29 = (32 if curve == "pallas" else 48), O.montgomery_params(P.p).n
            self.field_bits = O.log2(P.p)
        self.q = P.q
        self.gx, self.gy = P.gx, P.gy
        self.h = self.L.port_create(C.byref(pp), n29)
        self.point_words = self.L.port_point_words(self.h)

    def close(self):
        if self.h:
            self.L.port_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_window(self, n: int, form: int = 0) -> int:
        """The reference's window policy (src/msm-common.ts:8-57)."""
        lg = O.log2(n) if n > 1 else 0
        if self.curve == "ed-on-bls12-377" or form != 0:
            return O.window_size(self.field_bits, lg)
        return O.window_size_affine(self.field_bits, lg)

    def prepare_points(self, points_le: bytes, n: int, threads: int = 1):
        """LE bytes -> the reference's in-memory layout (untimed set-up, like Parallel.pointsFromBytes)."""
        buf = (C.c_uint32 * (max(n, 1) * self.point_words))()
        src = (C.c_uint8 * max(len(points_le), 1)).from_buffer_copy(points_le or b"\0")
        self.L.port_points_from_bytes(self.h, src, n, self.nbytes, buf, threads)
        return buf

    def msm(self, scalars_le: bytes, points, n: int, threads: int = 1, window_bits: int = 0, form: int = 0):
        """Returns (x, y, is_zero, seconds).  `points`: result of prepare_points."""
        c = window_bits or self.default_window(max(n, 1), form)
        ox = (C.c_uint8 * 64)()
        oy = (C.c_uint8 * 64)()
        sec = C.c_double()
        src = (C.c_uint8 * max(len(scalars_le), 1)).from_buffer_copy(scalars_le or b"\0")
        z = self.L.port_msm(self.h, src, points, n, threads, c, form, ox, oy, self.nbytes, C.byref(sec))
        return (int.from_bytes(bytes(ox)[: self.nbytes], "little"), int.from_bytes(bytes(oy)[: self.nbytes], "little"),
                bool(z), sec.value)

    # -- seeded inputs, byte-identical to the CUDA generators (msm_b200_random_points / _scalars)
    def random_scalars(self, n: int, seed: int, threads: int = 1, first: int = 0) -> bytes:
        out = (C.c_uint8 * (32 * max(n, 1)))()
        self.L.port_random_scalars(self.h, out, first, n, seed & (2**64 - 1), threads)
        return bytes(out)[: 32 * n]

    def random_points(self, n: int, seed: int, threads: int = 1, first: int = 0) -> bytes:
        out = (C.c_uint8 * (2 * self.nbytes * max(n, 1)))()
        gx = (C.c_uint8 * 64).from_buffer_copy(self.gx.to_bytes(64, "little"))
        gy = (C.c_uint8 * 64).from_buffer_copy(self.gy.to_bytes(64, "little"))
        self.L.port_random_points(self.h, gx, gy, out, first, n, seed & (2**64 - 1), self.nbytes, threads)
        return bytes(out)[: 2 * self.nbytes * n]

    def mul_ns(self, iters: int = 200000) -> float:
        return self.L.port_mul_ns(self.h, iters)
