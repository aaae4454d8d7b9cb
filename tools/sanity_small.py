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

# the shared-bucket paths (window tables need >= 2^14 points, 2^13 for twisted Edwards) and the multi-GPU layer on
# one device
for curve, n in (("bls12-377", (1 << 14) + 3), ("pallas", 1 << 14), ("ed-on-bls12-377", (1 << 13) + 5)):
    with mz.MultiMsmEngine(curve, [0]) as m:
        eng = m.shards[0]
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        d_pts = eng.dev_alloc(n * pb)
        d_sc = eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, n, 7 + n)
        eng.random_scalars_device(d_sc, n, 9 + n)
        m.set_bases_sharded([d_pts], [n])
        a = m.run_sharded([d_sc])
        assert a.timing["shared_buckets"] == 1, curve
        b = eng.run(d_sc, n, on_device=True, window_bits=11)
        assert (a.x, a.y) == (b.x, b.y), (curve, n)
        pts, sc = eng.d2h(d_pts, n * pb), eng.d2h(d_sc, n * 32)
        c = m.msm(sc, pts, n)
        assert (a.x, a.y) == (c.x, c.y), (curve, n)
    print(curve, "tables + multi ok", flush=True)
