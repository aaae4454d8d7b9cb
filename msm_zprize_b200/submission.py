"""Host-side mirror of the ZPrize-harness entry `compute_msm(inputPoints, inputScalars)`
(scripts/zprize23/submission-bls377.ts:20-65 for BLS12-377, scripts/zprize23/submission.ts:19-35 for
ed-on-bls12-377): same argument forms, same result shape, over one call of the C ABI
(`msm_b200_msm`, LE_BYTES layout -- the unpacking and the Montgomery conversion that the reference does
in `Parallel.pointsFromBytes / scalarsFromBytes` run on the GPU, csrc/kernels_weierstrass.cuh k_ingest_points).

    sub = Submission("bls12-377")
    sub.compute_msm(points_bytes, scalars_bytes)            # Buffer form: 96 B (64 B TE) per point, 32 B per scalar
    sub.compute_msm([{"x": .., "y": .., "isZero": False}], [s0])  # bigint form
    -> {"x": int, "y": int}

Differences from the reference, on purpose: every call uses the safe addition rules, so the reference's
`samePoints` branch (:44-57, msm vs msmUnsafe) needs no counterpart; there is no nMax = 2^20 limit.
"""
from __future__ import annotations

from typing import Sequence, Union

import numpy as np

from . import _lib as L
from .engine import FIELD_BYTES, MsmEngine

BytesLike = Union[bytes, bytearray, memoryview, np.ndarray]


def _is_bytes(a) -> bool:
    return isinstance(a, (bytes, bytearray, memoryview)) or (isinstance(a, np.ndarray) and a.dtype == np.uint8)


def _nbytes(a) -> int:
    return a.nbytes if isinstance(a, (np.ndarray, memoryview)) else len(a)


def scalars_to_bytes(scalars: Sequence) -> bytes:
    """bigint[] or Uint32Array[] (8 little-endian words each) -> n x 32 B little-endian
    (scalarsFromBigint, submission-bls377.ts:94-102; the u32 form of the harness type)."""
    out = bytearray()
    for s in scalars:
        if isinstance(s, (int, np.integer)):
            s = int(s)
            if s < 0 or s >> 256:
                raise L.MsmError(L.E_INVALID, "scalar out of range")
            out += s.to_bytes(32, "little")
        else:
            w = np.asarray(s, dtype=np.uint32)
            if w.shape != (8,):
                raise L.MsmError(L.E_INVALID, "u32 scalar must have 8 words")
            out += w.astype("<u4").tobytes()
    return bytes(out)


def points_to_bytes(points: Sequence, field_bytes: int):
    """BigIntPoint[] ({x, y[, isZero]}) or U32ArrayPoint[] -> (n' x (x | y) little-endian canonical bytes,
    kept indices).  The byte format has no encoding of the zero point (src/parallel.ts:107-108; the
    reference writes a flag byte instead, Affine.writeBigints src/curve-affine.ts:235-255), and a zero
    point contributes nothing to the sum, so such entries are dropped together with their scalars."""
    out = bytearray()
    keep = []
    words = field_bytes // 4
    for i, pt in enumerate(points):
        x, y = pt["x"], pt["y"]
        if isinstance(x, (int, np.integer)):
            if pt.get("isZero", False):
                continue
            for v in (int(x), int(y)):
                if v < 0 or v >> (8 * field_bytes):
                    raise L.MsmError(L.E_INVALID, "coordinate out of range")
                out += v.to_bytes(field_bytes, "little")
        else:
            for v in (x, y):
                w = np.asarray(v, dtype=np.uint32)
                if w.shape != (words,):
                    raise L.MsmError(L.E_INVALID, "u32 coordinate must have %d words" % words)
                out += w.astype("<u4").tobytes()
        keep.append(i)
    return bytes(out), keep


class Submission:
    """One curve's `compute_msm`; owns an engine (= the module-level state of the reference script)."""

    def __init__(self, curve: str = "bls12-377", device: int = 0):
        self.engine = MsmEngine(curve, device=device)
        self.field_bytes = FIELD_BYTES[self.engine.curve]

    def close(self):
        self.engine.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def compute_msm(self, inputPoints, inputScalars) -> dict:
        if _is_bytes(inputScalars):
            if _nbytes(inputScalars) % 32:
                raise L.MsmError(L.E_INVALID, "scalar buffer is not a multiple of 32 bytes")
            n = _nbytes(inputScalars) // 32
            sc = bytes(inputScalars) if not isinstance(inputScalars, np.ndarray) else inputScalars
        else:
            n = len(inputScalars)
            sc = scalars_to_bytes(inputScalars)
        if _is_bytes(inputPoints):
            pts = inputPoints
        else:
            if len(inputPoints) != n:
                raise L.MsmError(L.E_INVALID, "one point per scalar expected")
            pts, keep = points_to_bytes(inputPoints, self.field_bytes)
            if len(keep) != n:  # zero points dropped
                raw = np.frombuffer(bytes(sc), dtype=np.uint8).reshape(n, 32)
                sc = np.ascontiguousarray(raw[keep]).tobytes()
                n = len(keep)
        if _nbytes(pts) != n * 2 * self.field_bytes:
            raise L.MsmError(L.E_INVALID, "point buffer does not hold one point per scalar")
        if n == 0:
            zero_y = 1 if self.engine.curve == L.CURVE_ED_ON_BLS12_377 else 0
            return {"x": 0, "y": zero_y, "isZero": True}
        res = self.engine.msm(sc, pts, n)
        return {"x": res.x, "y": res.y, "isZero": res.is_zero}
