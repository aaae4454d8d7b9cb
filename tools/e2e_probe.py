#!/usr/bin/env python3
"""Phase timings of msm_b200_run with host (pinned) scalars against device-resident scalars (dev tool)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402
from msm_zprize_b200.engine import PinnedBuffer  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << lg
with mz.MsmEngine("bls12-377") as eng:
    pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
    d_pts, d_sc = eng.dev_alloc(n * pb), eng.dev_alloc(n * 32)
    eng.random_points_device(d_pts, n, 1)
    eng.random_scalars_device(d_sc, n, 2)
    eng.set_bases_device(d_pts, n)
    h = PinnedBuffer(n * 32)
    h.array[:] = eng.d2h(d_sc, n * 32)
    for name, fn in (("device scalars", lambda: eng.run(d_sc, n, on_device=True)), ("pinned host scalars", lambda: eng.run(h.array, n))):
        best = None
        for _ in range(10):
            t0 = time.perf_counter()
            r = fn()
            wall = (time.perf_counter() - t0) * 1e3
            if best is None or wall < best[0]:
                best = (wall, r.timing)
        t = best[1]
        print("%-20s wall %.3f ms | total %.3f h2d %.3f digits %.3f sort %.3f acc %.3f reduce %.3f d2h %.3f" % (
            name, best[0], t["total_ms"], t["h2d_ms"], t["digits_ms"], t["sort_ms"], t["accumulate_ms"], t["reduce_ms"], t["d2h_ms"]))
