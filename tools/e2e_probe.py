import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import msm_zprize_b200 as mz
from msm_zprize_b200.engine import PinnedBuffer
n = 1 << 18
eng = mz.MsmEngine("bls12-377")
d_pts = eng.dev_alloc(n * 96); d_sc = eng.dev_alloc(n * 32)
eng.random_points_device(d_pts, n, 1); eng.random_scalars_device(d_sc, n, 2)
hp = PinnedBuffer(n * 96); hs = PinnedBuffer(n * 32)
hp.array[:] = eng.d2h(d_pts, n * 96); hs.array[:] = eng.d2h(d_sc, n * 32)
for it in range(4):
    t0 = time.perf_counter(); r = eng.msm(hs.array, hp.array, n); t1 = time.perf_counter()
    print("msm e2e ms", round((t1 - t0) * 1e3, 3), {k: round(v, 3) for k, v in r.timing.items() if k.endswith("_ms")})
# raw copy speed
for it in range(3):
    t0 = time.perf_counter(); eng.h2d(d_pts, hp.array); t1 = time.perf_counter()
    print("h2d 25MB ms", round((t1 - t0) * 1e3, 3), "GB/s", round(n * 96 / (t1 - t0) / 1e9, 1))
eng.set_bases_device(d_pts, n)
for it in range(3):
    t0 = time.perf_counter(); r = eng.run(hs.array, n); t1 = time.perf_counter()
    print("run (host scalars) ms", round((t1 - t0) * 1e3, 3))
    t0 = time.perf_counter(); r = eng.run(d_sc, n, on_device=True); t1 = time.perf_counter()
    print("run (device scalars) ms", round((t1 - t0) * 1e3, 3))
