#!/usr/bin/env python3
"""Small MSMs on every curve / form (used under compute-sanitizer memcheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402

for curve in ("bls12-377", "pallas", "bls12-381", "ed-on-bls12-377"):
    with mz.MsmEngine(curve) as eng:
        for n in (1, 3, 64, 1000, 5000):
            pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
            d_pts = eng.dev_alloc(n * pb)
            d_sc = eng.dev_alloc(n * 32)
            eng.random_points_device(d_pts, n, 7 + n)
            eng.random_scalars_device(d_sc, n, 9 + n)
            eng.set_bases_device(d_pts, n)
            a = eng.run(d_sc, n, on_device=True)
            b = eng.run(d_sc, n, on_device=True, window_bits=5)
            assert (a.x, a.y) == (b.x, b.y), (curve, n)
            if curve != "ed-on-bls12-377":
                p = eng.run(d_sc, n, on_device=True, form=mz.FORM_PROJECTIVE)
                assert (a.x, a.y) == (p.x, p.y), (curve, n)
            eng.dev_free(d_pts)
            eng.dev_free(d_sc)
    print(curve, "ok", flush=True)
