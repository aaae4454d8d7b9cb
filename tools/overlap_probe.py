#!/usr/bin/env python3
"""Dev probe: do two independent MSMs on two streams of one GPU overlap?  Two engines (own stream each)
run n-point MSMs concurrently from two host threads; compared with one engine alone and with one engine on 2n."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 17
n = 1 << lg


def setup(eng, n, seed):
    pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
    d_pts = eng.dev_alloc(n * pb)
    d_sc = eng.dev_alloc(n * 32)
    eng.random_points_device(d_pts, n, 0xB200 + seed)
    eng.random_scalars_device(d_sc, n, 0x5CA1A + seed)
    eng.set_bases_device(d_pts, n)
    return d_sc


engs = [mz.MsmEngine("bls12-377") for _ in range(2)]
scs = [setup(e, n, i) for i, e in enumerate(engs)]
big = mz.MsmEngine("bls12-377")
sc_big = setup(big, 2 * n, 7)


def run(e, sc, n, reps, out, i, barrier):
    for _ in range(3):
        e.run(sc, n, on_device=True)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        e.run(sc, n, on_device=True)
    out[i] = (time.perf_counter() - t0) / reps * 1e3


reps = 20
out = [0, 0]
b = threading.Barrier(1)
run(engs[0], scs[0], n, reps, out, 0, b)
print(f"one engine alone, 2^{lg}: {out[0]:.3f} ms per MSM")
b = threading.Barrier(1)
run(big, sc_big, 2 * n, reps, out, 0, b)
print(f"one engine alone, 2^{lg + 1}: {out[0]:.3f} ms per MSM")
b = threading.Barrier(2)
ths = [threading.Thread(target=run, args=(engs[i], scs[i], n, reps, out, i, b)) for i in range(2)]
for t in ths:
    t.start()
for t in ths:
    t.join()
print(f"two engines concurrently, 2^{lg} each: {out[0]:.3f} / {out[1]:.3f} ms per MSM pair")
