"""Host-side mirror of the reference's `Parallel` object (src/parallel.ts:135-145,251-259): the same
function names and argument meaning over the CUDA engine, so that tests read like the reference's
(src/msm.test.ts, scripts/msm-weierstrass.ts).

    BLS12377 = create_weierstrass("bls12-377")
    points = BLS12377.Parallel.randomPointsFast(N)        # device buffer (seeded; the reference is unseeded)
    scalars = BLS12377.Parallel.randomScalars(N)
    out = BLS12377.Parallel.msmUnsafe(scalars, points, N, True)   # {"result": MsmResult, "log": [...]}
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass, field
from typing import Optional

from . import _lib as L
from .engine import MsmEngine


_uids = itertools.count(1)


@dataclass
class DeviceBuffer:
    """What a wasm-memory pointer is in the reference: a handle to inputs that already live where the
    MSM runs.  `layout` is LE_BYTES for generated / converted data.

    `uid` is unique per allocation and `version` counts the writes into it: the resident-bases cache of
    `Parallel` is keyed on both, never on the address (cudaMalloc hands a freed address out again, and a
    buffer can be regenerated in place)."""
    ptr: int
    n: int
    layout: int = L.LAYOUT_LE_BYTES
    uid: int = field(default_factory=lambda: next(_uids))
    version: int = 0

    def touch(self):
        """call after writing new contents through `ptr` by other means than this module"""
        self.version += 1


class Parallel:
    def __init__(self, engine: MsmEngine, twisted_edwards: bool):
        self.engine = engine
        self.te = twisted_edwards
        self._bases_key = None  # (uid, version) of the DeviceBuffer whose points are resident
        self._seed = 0xB200

    # -- input helpers (src/curve-random.ts:24-91,151-194; src/parallel.ts:97-133,209-249)
    def randomPointsFast(self, n: int, seed: Optional[int] = None) -> DeviceBuffer:
        ptr = self.engine.dev_alloc(n * self.engine.point_bytes(L.LAYOUT_LE_BYTES))
        self.engine.random_points_device(ptr, n, self._next_seed(seed))
        return DeviceBuffer(ptr, n)

    def randomScalars(self, n: int, seed: Optional[int] = None) -> DeviceBuffer:
        ptr = self.engine.dev_alloc(n * 32)
        self.engine.random_scalars_device(ptr, n, self._next_seed(seed))
        return DeviceBuffer(ptr, n)

    def pointsFromBytes(self, data: bytes, n: int) -> DeviceBuffer:
        ptr = self.engine.dev_alloc(max(len(data), 16))
        self.engine.h2d(ptr, data)
        return DeviceBuffer(ptr, n)

    def scalarsFromBytes(self, data: bytes, n: int) -> DeviceBuffer:
        ptr = self.engine.dev_alloc(max(len(data), 16))
        self.engine.h2d(ptr, data)
        return DeviceBuffer(ptr, n)

    def regeneratePointsFast(self, buf: DeviceBuffer, seed: Optional[int] = None) -> DeviceBuffer:
        """randomPointsFast into an existing buffer (the reference regenerates at the same wasm offset after a
        scope reset, scripts/msm-weierstrass.ts:12-22)."""
        self.engine.random_points_device(buf.ptr, buf.n, self._next_seed(seed))
        buf.touch()
        return buf

    def free(self, buf: DeviceBuffer):
        if self._bases_key is not None and self._bases_key[0] == buf.uid:
            self._bases_key = None
        self.engine.dev_free(buf.ptr)
        buf.ptr, buf.n = 0, 0

    def _next_seed(self, seed):
        if seed is not None:
            return seed
        self._seed += 1
        return self._seed

    # -- the MSM entry points (src/msm-batched-affine.ts:74-83,573-587; src/parallel.ts:69-87; src/msm-basic.ts:34-43)
    def _msm(self, scalars: DeviceBuffer, points: DeviceBuffer, N: int, verbose: bool, options, form):
        options = options or {}
        if N > points.n or N > scalars.n:
            raise L.MsmError(L.E_INVALID, "N exceeds the input buffers")
        if not points.ptr:
            raise L.MsmError(L.E_INVALID, "points buffer was freed")
        if self._bases_key != (points.uid, points.version):
            self.engine.set_bases_device(points.ptr, points.n, points.layout)
            self._bases_key = (points.uid, points.version)
        res = self.engine.run(scalars.ptr, N, layout=scalars.layout, form=form,
                              window_bits=int(options.get("c", 0) or 0), on_device=True)
        log = [[f"{k[:-3]}... {v:.2f}ms"] for k, v in res.timing.items() if k.endswith("_ms")] if verbose else []
        return {"result": res, "log": log}

    def msm(self, scalars, points, N, verbose=False, options=None):
        return self._msm(scalars, points, N, verbose, options,
                         L.FORM_TE_EXTENDED if self.te else L.FORM_AFFINE_GLV)

    def msmUnsafe(self, scalars, points, N, verbose=False, options=None):
        # the engine always uses the safe addition rules (src/curve-affine.ts:376-458); identical result
        return self.msm(scalars, points, N, verbose, options)

    def msmProjective(self, scalars, points, N, options=None):
        if self.te:
            raise L.MsmError(L.E_INVALID, "msmProjective is a Weierstrass entry point")
        return self._msm(scalars, points, N, False, options, L.FORM_PROJECTIVE)


class CurveBundle:
    def __init__(self, name: str, device: int = 0):
        self.name = name
        self.engine = MsmEngine(name, device=device)
        self.Parallel = Parallel(self.engine, name == "ed-on-bls12-377")

    def close(self):
        self.engine.close()


def create_weierstrass(name: str, device: int = 0) -> CurveBundle:
    """Weierstrass.create(params) (src/parallel.ts:40-177) for "bls12-377", "pallas" or "bls12-381"."""
    assert name in ("bls12-377", "pallas", "bls12-381")
    return CurveBundle(name, device)


def create_twisted_edwards(name: str = "ed-on-bls12-377", device: int = 0) -> CurveBundle:
    """TwistedEdwards.create(params) (src/parallel.ts:179-289)."""
    assert name == "ed-on-bls12-377"
    return CurveBundle(name, device)
