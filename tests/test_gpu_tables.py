"""Shared-bucket mode: set_bases builds the window tables 2^(kc) G of the resident point set and the MSM then
adds the digits of all windows into ONE set of buckets.  Same group element as the classic layout, so every
case is checked against the CPU port (and against the classic path of the same engine: an explicit window
size other than the tables' one, the one-shot call, and an engine created with MSM_B200_TABLES=0)."""
import os
import random

import numpy as np
import pytest

from oracle import bigint_oracle as O
from tests import inputs as I

pytestmark = pytest.mark.gpu
PARAMS = {"bls12-377": O.BLS12_377, "pallas": O.PALLAS, "bls12-381": O.BLS12_381}


@pytest.fixture(scope="module")
def mz():
    import msm_zprize_b200 as m
    return m


def _port_msm(name, sc, pts, n):
    from oracle.port import Port
    port = Port(name)
    t = os.cpu_count() or 1
    return port.msm(sc, port.prepare_points(pts, n, t), n, t)[:3]


@pytest.mark.parametrize("name", ["bls12-377", "pallas", "bls12-381"])
def test_shared_buckets_equal_classic_and_port(mz, name):
    from oracle.port import Port
    n = (1 << 14) + 321
    port = Port(name)
    pts = port.random_points(n, 0x7AB1E, 8)
    sc = port.random_scalars(n, 0x7AB1F, 8)
    want = _port_msm(name, sc, pts, n)
    with mz.MsmEngine(name) as eng:
        eng.set_bases(pts, n)
        a = eng.run(sc, n)
        assert a.timing["shared_buckets"] == 1 and a.timing["window_bits"] == 16
        assert (a.x, a.y, a.is_zero) == want
        b = eng.run(sc, n, window_bits=15)  # not the tables' window: classic layout over the same bases
        assert b.timing["shared_buckets"] == 0 and (b.x, b.y, b.is_zero) == want
        one = eng.msm(sc, pts, n)  # one-shot: never builds tables
        assert one.timing["shared_buckets"] == 0 and (one.x, one.y, one.is_zero) == want
        # a prefix of the resident set (the table stride stays the resident count)
        eng.set_bases(pts, n)
        k = (1 << 14) + 5
        p = eng.run(sc, k)
        assert p.timing["shared_buckets"] == 1
        assert (p.x, p.y, p.is_zero) == _port_msm(name, sc[:32 * k], pts[:k * 2 * port.nbytes], k)
        # projective form over the same resident bases ignores the tables
        q = eng.run(sc, n, form=mz.FORM_PROJECTIVE)
        assert (q.x, q.y) == want[:2]
    os.environ["MSM_B200_TABLES"] = "0"
    try:
        with mz.MsmEngine(name) as eng:
            eng.set_bases(pts, n)
            c = eng.run(sc, n)
            assert c.timing["shared_buckets"] == 0 and (c.x, c.y, c.is_zero) == want
    finally:
        del os.environ["MSM_B200_TABLES"]


def test_shared_buckets_safe_additions(mz):
    """batchAddNew semantics (src/curve-affine.ts:376-458) at a size that uses the tables: repeated points
    (doublings), P / -P pairs (cancellation), scalars 0 / 1 / q - 1, and -- in the reference's in-memory
    layout -- points whose isNonZero flag is 0 (their table entries must stay the point at infinity)."""
    name, params = "bls12-377", O.BLS12_377
    aff = O.WeierstrassAffine(params)
    q = params.q
    from oracle.port import Port
    port = Port(name)
    n = 1 << 14
    base = port.random_points(n, 0x5AFE, 8)
    P = [(int.from_bytes(base[96 * i:96 * i + 48], "little"), int.from_bytes(base[96 * i + 48:96 * i + 96], "little"))
         for i in range(n)]
    rng = random.Random(5)
    sc = [rng.randrange(q) for _ in range(n)]
    # blocks of special structure inside an otherwise random instance
    for i in range(0, 64):
        P[i] = P[0]  # 64 copies of one point
        sc[i] = sc[0]  # ... with one scalar: every digit meets its copies in the same bucket
    for i in range(100, 200, 2):
        P[i + 1] = aff.negate(P[i])  # P, -P with equal scalars cancel
        sc[i + 1] = sc[i]
    sc[300], sc[301], sc[302] = 0, 1, q - 1
    pts_le, sc_le = I.points_le(P, 48), I.scalars_le(sc)
    want = _port_msm(name, sc_le, pts_le, n)
    with mz.MsmEngine(name) as eng:
        eng.set_bases(pts_le, n)
        r = eng.run(sc_le, n)
        assert r.timing["shared_buckets"] == 1 and (r.x, r.y, r.is_zero) == want
        # zero points (flag 0) at every third index: they contribute nothing
        Pz = [None if i % 3 == 0 else P[i] for i in range(n)]
        keep = [i for i in range(n) if i % 3]
        want_z = _port_msm(name, I.scalars_le([sc[i] for i in keep]), I.points_le([P[i] for i in keep], 48), len(keep))
        eng.set_bases(I.points_limb29(Pz, params.p), n, mz.LAYOUT_LIMB29_MONT)
        z = eng.run(I.scalars_limb29(sc, q), n, mz.LAYOUT_LIMB29_MONT)
        assert z.timing["shared_buckets"] == 1 and (z.x, z.y, z.is_zero) == want_z


def test_shared_buckets_skewed_scalars(mz):
    """All scalars equal: each window's digit lands in one bucket, so a handful of buckets hold n entries each
    (deep pairwise tree, more rounds than the first scan covers); all scalars zero: nothing to add."""
    name, params = "pallas", O.PALLAS
    aff = O.WeierstrassAffine(params)
    from oracle.port import Port
    port = Port(name)
    n = 1 << 14
    pts = port.random_points(n, 0x5CE, 8)
    s = 0x1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF1234567890ABCDEF % params.q
    sc = I.scalars_le([s] * n)
    with mz.MsmEngine(name) as eng:
        eng.set_bases(pts, n)
        r = eng.run(sc, n)
        assert r.timing["shared_buckets"] == 1
        assert (r.x, r.y, r.is_zero) == _port_msm(name, sc, pts, n)
        z = eng.run(bytes(32 * n), n)
        assert z.is_zero


@pytest.mark.parametrize("lg,c_tab", [(13, 14), (17, 16)])
def test_te_shared_buckets_equal_classic_and_port(mz, lg, c_tab):
    """ed-on-bls12-377 (generic bucket method, src/msm-basic.ts): resident bases get tables 2^(kc) P in the cached
    (y+x, y-x, 2dxy) form and the window size that goes with them."""
    from oracle.port import Port
    name = "ed-on-bls12-377"
    n = (1 << lg) - 77
    port = Port(name)
    pts = port.random_points(n, 0x7E0 + lg, 8)
    sc = port.random_scalars(n, 0x7E1 + lg, 8)
    want = _port_msm(name, sc, pts, n)
    with mz.MsmEngine(name) as eng:
        eng.set_bases(pts, n)
        a = eng.run(sc, n)
        assert a.timing["shared_buckets"] == 1 and a.timing["window_bits"] == c_tab
        assert (a.x, a.y, a.is_zero) == want
        b = eng.run(sc, n, window_bits=11)  # classic layout over the same resident bases
        assert b.timing["shared_buckets"] == 0 and (b.x, b.y, b.is_zero) == want
        one = eng.msm(sc, pts, n)
        assert one.timing["shared_buckets"] == 0 and (one.x, one.y, one.is_zero) == want
        eng.set_bases(pts, n)
        k = n - 1000
        p = eng.run(sc, k)
        assert p.timing["shared_buckets"] == 1 and (p.x, p.y, p.is_zero) == _port_msm(name, sc[:32 * k], pts[:64 * k], k)


def test_te_shared_buckets_special_inputs(mz):
    """Unified additions under the tables: repeated points, P / -P pairs, the neutral point (0, 1) as an input,
    scalars 0 / 1 / q - 1, and all-equal scalars (a few buckets hold everything: virtual-bucket split + tree)."""
    name = "ed-on-bls12-377"
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    from oracle.port import Port
    port = Port(name)
    n = 1 << 13
    base = port.random_points(n, 0x7ED, 8)
    P = [(int.from_bytes(base[64 * i:64 * i + 32], "little"), int.from_bytes(base[64 * i + 32:64 * i + 64], "little"))
         for i in range(n)]
    rng = random.Random(7)
    sc = [rng.randrange(te.q) for _ in range(n)]
    for i in range(64):
        P[i], sc[i] = P[0], sc[0]
    for i in range(100, 200, 2):
        P[i + 1] = te.to_affine(te.negate(te.from_affine(P[i])))
        sc[i + 1] = sc[i]
    P[250] = (0, 1)
    sc[300], sc[301], sc[302] = 0, 1, te.q - 1
    pts_le, sc_le = I.points_le(P, 32), I.scalars_le(sc)
    with mz.MsmEngine(name) as eng:
        eng.set_bases(pts_le, n)
        r = eng.run(sc_le, n)
        assert r.timing["shared_buckets"] == 1 and (r.x, r.y, r.is_zero) == _port_msm(name, sc_le, pts_le, n)
        same = I.scalars_le([sc[5]] * n)
        r = eng.run(same, n)
        assert r.timing["shared_buckets"] == 1 and (r.x, r.y, r.is_zero) == _port_msm(name, same, pts_le, n)
        z = eng.run(bytes(32 * n), n)
        assert z.is_zero and (z.x, z.y) == (0, 1)
