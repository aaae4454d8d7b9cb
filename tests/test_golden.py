"""Committed golden vectors (tests/golden/msm_vectors.json, made by tools/make_golden.py):
the oracle, the C++ port and -- on the GPU box -- the CUDA path must all reproduce them."""
import json
import os

import pytest

from oracle import bigint_oracle as O
from tests import inputs as I

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "msm_vectors.json")))
NB = {"bls12-377": 48, "pallas": 32, "ed-on-bls12-377": 32}


def _case(c):
    pts = [(int(x, 16), int(y, 16)) for x, y in c["points"]]
    sc = [int(s, 16) for s in c["scalars"]]
    return pts, sc, (int(c["result"][0], 16), int(c["result"][1], 16))


@pytest.mark.parametrize("idx", range(len(GOLD["cases"])))
def test_oracle_and_port_reproduce_golden(idx):
    from oracle.port import Port
    c = GOLD["cases"][idx]
    pts, sc, want = _case(c)
    if c["curve"] == "ed-on-bls12-377":
        te = O.TwistedEdwards(O.ED_ON_BLS12_377)
        assert te.to_affine(O.msm_naive(te, sc, [te.from_affine(p) for p in pts])) == want
    else:
        aff = O.WeierstrassAffine(O.BLS12_377 if c["curve"] == "bls12-377" else O.PALLAS)
        assert O.msm_naive(aff, sc, pts) == want
    port = Port(c["curve"])
    prep = port.prepare_points(I.points_le(pts, NB[c["curve"]]), c["n"])
    x, y, z, _ = port.msm(I.scalars_le(sc), prep, c["n"], 2, 4)
    assert (x, y) == want


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(len(GOLD["cases"])))
def test_gpu_reproduces_golden(idx):
    import msm_zprize_b200 as mz
    c = GOLD["cases"][idx]
    pts, sc, want = _case(c)
    with mz.MsmEngine(c["curve"]) as eng:
        r = eng.msm(I.scalars_le(sc), I.points_le(pts, NB[c["curve"]]), c["n"])
        assert (r.x, r.y) == want
        if c["curve"] != "ed-on-bls12-377":
            r = eng.msm(I.scalars_le(sc), I.points_le(pts, NB[c["curve"]]), c["n"], form=mz.FORM_PROJECTIVE)
            assert (r.x, r.y) == want


@pytest.mark.gpu
def test_gpu_kats_from_golden_file():
    import msm_zprize_b200 as mz
    for curve, k in GOLD["kat"].items():
        P = (int(k["point"][0], 16), int(k["point"][1], 16))
        sc = [int(s, 16) for s in k["scalars"]]
        with mz.MsmEngine(curve) as eng:
            r = eng.msm(I.scalars_le(sc), I.points_le([P, P], NB[curve]), 2)
        assert (r.x, r.y) == P


@pytest.mark.gpu
def test_parallel_mirror_reads_like_the_reference():
    # scripts/msm-weierstrass.ts:53-110 (runMsm) through the Parallel mirror
    from msm_zprize_b200.parallel import create_twisted_edwards, create_weierstrass
    B = create_weierstrass("bls12-377")
    N = 1 << 10
    points = B.Parallel.randomPointsFast(N)
    scalars = B.Parallel.randomScalars(N)
    a = B.Parallel.msmUnsafe(scalars, points, N, True)
    b = B.Parallel.msm(scalars, points, N)
    p = B.Parallel.msmProjective(scalars, points, N, {"c": 7})
    assert (a["result"].x, a["result"].y) == (b["result"].x, b["result"].y) == (p["result"].x, p["result"].y)
    assert any("accumulate" in row[0] for row in a["log"])
    # against the oracle on the same inputs
    raw_p = B.engine.d2h(points.ptr, N * 96).tobytes()
    raw_s = B.engine.d2h(scalars.ptr, N * 32).tobytes()
    aff = O.WeierstrassAffine(O.BLS12_377)
    P = [(int.from_bytes(raw_p[i * 96:i * 96 + 48], "little"), int.from_bytes(raw_p[i * 96 + 48:(i + 1) * 96], "little")) for i in range(N)]
    S = [int.from_bytes(raw_s[i * 32:(i + 1) * 32], "little") for i in range(N)]
    assert (a["result"].x, a["result"].y) == O.msm(aff, S, P)
    B.close()
    T = create_twisted_edwards()
    pts = T.Parallel.randomPointsFast(256)
    sc = T.Parallel.randomScalars(256)
    r = T.Parallel.msm(sc, pts, 256)["result"]
    assert not r.is_zero
    T.close()


@pytest.mark.gpu
def test_parallel_bases_cache_follows_contents_not_addresses():
    """free -> allocate again (cudaMalloc hands the same address out) -> other points: the mirror must not
    answer with the MSM of the old points; the same for points regenerated in place."""
    import msm_zprize_b200.parallel as P
    B = P.create_weierstrass("pallas")
    N = 512
    sc = B.Parallel.randomScalars(N, seed=5)
    p1 = B.Parallel.randomPointsFast(N, seed=1)
    a = B.Parallel.msm(sc, p1, N)["result"]
    B.Parallel.free(p1)
    p2 = B.Parallel.randomPointsFast(N, seed=2)  # very likely the same device address
    b = B.Parallel.msm(sc, p2, N)["result"]
    fresh = P.create_weierstrass("pallas")
    sc2 = fresh.Parallel.randomScalars(N, seed=5)
    want = fresh.Parallel.msm(sc2, fresh.Parallel.randomPointsFast(N, seed=2), N)["result"]
    assert (b.x, b.y) == (want.x, want.y) and (a.x, a.y) != (b.x, b.y)
    B.Parallel.regeneratePointsFast(p2, seed=1)
    c = B.Parallel.msm(sc, p2, N)["result"]
    assert (c.x, c.y) == (a.x, a.y)
    B.close()
    fresh.close()
