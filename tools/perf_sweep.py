#!/usr/bin/env python3
"""Phase timings of the MSM over a range of sizes with device-generated inputs (dev tool)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--curve", default="bls12-377")
ap.add_argument("--sizes", default="16,18,20,22")
ap.add_argument("--windows", default="0")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--form", type=int, default=-1)
args = ap.parse_args()

eng = mz.MsmEngine(args.curve)
form = None if args.form < 0 else args.form
for lg in [int(s) for s in args.sizes.split(",")]:
    n = 1 << lg
    pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
    d_pts = eng.dev_alloc(n * pb)
    d_sc = eng.dev_alloc(n * 32)
    t0 = time.time()
    eng.random_points_device(d_pts, n, 0xB200 + lg)
    eng.random_scalars_device(d_sc, n, 0x5CA1A + lg)
    t1 = time.time()
    eng.set_bases_device(d_pts, n)
    t2 = time.time()
    print(f"# n=2^{lg}: gen {t1 - t0:.3f}s ingest {t2 - t1:.3f}s", flush=True)
    ref = None
    for c in [int(x) for x in args.windows.split(",")]:
        best = None
        for rep in range(args.reps):
            r = eng.run(d_sc, n, form=form, window_bits=c, on_device=True)
            if best is None or r.timing["total_ms"] < best.timing["total_ms"]:
                best = r
        t = best.timing
        if ref is None:
            ref = (best.x, best.y)
        ok = (best.x, best.y) == ref
        print(json.dumps({"lg": lg, "c": t["window_bits"], "K": t["n_windows"], "rounds": t["rounds"],
                          "total_ms": round(t["total_ms"], 3), "digits": round(t["digits_ms"], 3),
                          "sort": round(t["sort_ms"], 3), "acc": round(t["accumulate_ms"], 3),
                          "hot": round(t["hot_kernel_ms"], 3), "reduce": round(t["reduce_ms"], 3),
                          "launches": t["kernel_launches"], "shared": t["shared_buckets"], "Mpts/s": round(n / t["total_ms"] / 1e3, 2),
                          "same_as_first_c": ok}), flush=True)
    eng.dev_free(d_pts)
    eng.dev_free(d_sc)
