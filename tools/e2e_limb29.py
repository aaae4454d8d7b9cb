#!/usr/bin/env python3
"""Dev probe: end-to-end latency of msm_b200_msm() when the host buffers are in the REFERENCE'S in-memory format
(29-bit limbs, Montgomery R = 2^406, x | y | flag for points, 9 x 29-bit limbs for scalars: what the N-API shim
hands over from wasm memory), next to the little-endian byte format that bench.py times.  BLS12-377, n = 2^lg."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402
from msm_zprize_b200.engine import PinnedBuffer  # noqa: E402

P = 0x01ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
N29, W = 14, 29
R29 = pow(2, N29 * W, P)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << lg

eng = mz.MsmEngine("bls12-377")
d_pts = eng.dev_alloc(n * 96)
d_sc = eng.dev_alloc(n * 32)
eng.random_points_device(d_pts, n, 0xB200 + lg)
eng.random_scalars_device(d_sc, n, 0x5CA1A + lg)
pts = eng.d2h(d_pts, n * 96)
sc = eng.d2h(d_sc, n * 32)

t0 = time.time()
mask = (1 << W) - 1
pl = np.zeros((n, 2 * N29 + 1), dtype=np.uint32)
raw = pts.tobytes()
for i in range(n):
    for c in range(2):
        v = int.from_bytes(raw[96 * i + 48 * c: 96 * i + 48 * c + 48], "little") * R29 % P
        for k in range(N29):
            pl[i, c * N29 + k] = (v >> (W * k)) & mask
    pl[i, 2 * N29] = 1  # isNonZero flag byte (+3 pad)
sl = np.zeros((n, 9), dtype=np.uint32)
rs = sc.tobytes()
for i in range(n):
    v = int.from_bytes(rs[32 * i: 32 * i + 32], "little")
    for k in range(9):
        sl[i, k] = (v >> (W * k)) & mask
print(f"host conversion to the limb29 layouts: {time.time() - t0:.1f} s", flush=True)

h_pl, h_sl = PinnedBuffer(pl.nbytes), PinnedBuffer(sl.nbytes)
h_pl.array[:] = pl.view(np.uint8).reshape(-1)
h_sl.array[:] = sl.view(np.uint8).reshape(-1)
h_pb, h_sb = PinnedBuffer(n * 96), PinnedBuffer(n * 32)
h_pb.array[:] = pts
h_sb.array[:] = sc


def bench(f, reps=20):
    for _ in range(3):
        r = f()
    t = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t) / reps * 1e3, r


ms_b, rb = bench(lambda: eng.msm(h_sb.array, h_pb.array, n))
ms_l, rl = bench(lambda: eng.msm(h_sl.array, h_pl.array, n, mz.LAYOUT_LIMB29_MONT, mz.LAYOUT_LIMB29_MONT))
assert (rb.x, rb.y) == (rl.x, rl.y), "layouts disagree"
print(f"n=2^{lg}: LE bytes {ms_b:.3f} ms ({(n * 128) >> 20} MiB H2D), limb29 {ms_l:.3f} ms "
      f"({(pl.nbytes + sl.nbytes) >> 20} MiB H2D), same result")
