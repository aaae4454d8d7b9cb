#!/usr/bin/env python3
"""Randomised differential test: CUDA engine against the CPU port of the reference algorithm, for a given number
of seconds.  Random curve, size, window, layout path (resident with tables / classic / one-shot / projective),
with structured inputs mixed in (repeated points, P / -P pairs, zero scalars, small scalars, equal scalars).
Prints every mismatch and exits non-zero if there was one.  dev tool; the checker is oracle/ (test infrastructure)."""
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402
from oracle.port import Port  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = random.Random(seed)
threads = os.cpu_count() or 1
CURVES = ["bls12-377", "pallas", "bls12-381", "ed-on-bls12-377"]
ports = {c: Port(c) for c in CURVES}
engines = {c: mz.MsmEngine(c) for c in CURVES}
t_end = time.time() + seconds
cases = bad = 0
while time.time() < t_end:
    curve = rng.choice(CURVES)
    port, eng = ports[curve], engines[curve]
    te = curve == "ed-on-bls12-377"
    lg = rng.choice([0, 3, 6, 9, 11, 12, 13, 14, 14, 15, 15, 16] + ([17, 17, 18] if os.environ.get("FUZZ_LARGE") else []))
    n = max(1, rng.randrange(1 << lg, (2 << lg)))
    nb = port.nbytes
    pts = bytearray(port.random_points(n, rng.getrandbits(40), threads))
    sc = bytearray(port.random_scalars(n, rng.getrandbits(40), threads))
    # structure
    kind = rng.choice(["plain", "plain", "repeat", "negpairs", "zeros", "small", "equal", "mixed"])
    def setpt(i, j):
        pts[2 * nb * i:2 * nb * (i + 1)] = pts[2 * nb * j:2 * nb * (j + 1)]
    def setsc(i, v):
        sc[32 * i:32 * (i + 1)] = int(v).to_bytes(32, "little")
    if kind in ("repeat", "mixed") and n > 4:
        for i in range(1, min(n, rng.randrange(2, 200))):
            setpt(i, 0)
            if rng.random() < 0.5:
                sc[32 * i:32 * (i + 1)] = sc[0:32]
    if kind in ("negpairs", "mixed") and n > 8 and not te:
        p = {"bls12-377": 0, "pallas": 0, "bls12-381": 0}
        from oracle import bigint_oracle as O
        P = {"bls12-377": O.BLS12_377, "pallas": O.PALLAS, "bls12-381": O.BLS12_381}[curve].p
        for i in range(4, min(n - 1, 60), 2):
            setpt(i + 1, i)
            y = int.from_bytes(pts[2 * nb * (i + 1) + nb:2 * nb * (i + 2)], "little")
            pts[2 * nb * (i + 1) + nb:2 * nb * (i + 2)] = ((P - y) % P).to_bytes(nb, "little")
            sc[32 * (i + 1):32 * (i + 2)] = sc[32 * i:32 * (i + 1)]
    if kind in ("zeros", "mixed"):
        for i in rng.sample(range(n), min(n, rng.randrange(1, 50))):
            setsc(i, 0)
    if kind == "small":
        for i in range(n):
            setsc(i, rng.randrange(0, 1 << rng.choice([1, 8, 16, 64])))
    if kind == "equal":
        v = int.from_bytes(sc[0:32], "little")
        for i in range(n):
            setsc(i, v)
    if kind == "mixed" and n > 3:
        setsc(n - 1, port.q - 1)
        setsc(n - 2, 1)
    pts, sc = bytes(pts), bytes(sc)
    want = port.msm(sc, port.prepare_points(pts, n, threads), n, threads)[:3]
    path = rng.choice(["resident", "resident", "window", "oneshot", "projective"])
    c = 0
    if path == "resident":
        eng.set_bases(pts, n)
        r = eng.run(sc, n)
    elif path == "window":
        c = rng.randrange(1, 20)
        eng.set_bases(pts, n)
        r = eng.run(sc, n, window_bits=c)
    elif path == "oneshot":
        r = eng.msm(sc, pts, n)
    else:
        eng.set_bases(pts, n)
        r = eng.run(sc, n, form=None if te else mz.FORM_PROJECTIVE)
    cases += 1
    if (r.x, r.y, r.is_zero) != want:
        bad += 1
        print("MISMATCH", curve, "n", n, "kind", kind, "path", path, "c", c, "shared", r.timing["shared_buckets"], flush=True)
print("fuzz: %d cases, %d mismatches, seed %d, %.0f s" % (cases, bad, seed, seconds))
sys.exit(1 if bad else 0)
