"""Host-side mirror of the reference's `Parallel` object (src/parallel.ts:135-145,251-259): the same
function names and argument meaning over the CUDA engine, so that tests read like the reference's
(src/msm.test.ts, scripts/msm-weierstrass.ts).

    BLS12377 = create_weierstrass("bls12-377")
    points = BLS12377.Parallel.randomPointsFast(N)        # device buffer (seeded; the reference is unseeded)
    scalars = BLS12377.Parallel.randomScalars(N)
    out = BLS12377.Parallel.msmUnsafe(scalars, points, N, True)   # {"result": MsmResult, "log": [...]}
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

from . import _lib as L
from .engine import MsmEngine


@dataclass
class DeviceBuffer:
    """What a wasm-memory pointer is in the reference: a handle to inputs that already live where the
    MSM runs.  `layout` is LE_BYTES for generated / converted data."""
    ptr: int
    n: int
    layout: int = L.LAYOUT_LE_BYTES


class Parallel:
    def __init__(self, engine: MsmEngine, twisted_edwards: bool):
        self.engine = engine
        self.te = twisted_edwards
        self._bases: Optional[int] = None
        self._bases_n = 0
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

    def free(self, buf: DeviceBuffer):
        self.engine.dev_free(buf.ptr)

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
        if self._bases != points.ptr or self._bases_n < N:
            self.engine.set_bases_device(points.ptr, points.n, points.layout)
            self._bases, self._bases_n = points.ptr, points.n
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
