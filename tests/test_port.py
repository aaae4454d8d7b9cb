"""Pins the C++ CPU port (oracle/msm_port.cpp) to the python oracle.  CPU only."""
import pytest

from oracle import bigint_oracle as O
from oracle.port import Port
from tests import inputs as I


@pytest.mark.parametrize("name,params", [("bls12-377", O.BLS12_377), ("pallas", O.PALLAS), ("bls12-381", O.BLS12_381)])
@pytest.mark.parametrize("n,threads,c", [(1, 1, 0), (2, 1, 3), (17, 3, 4), (64, 4, 0), (300, 8, 6)])
def test_port_weierstrass(name, params, n, threads, c):
    aff = O.WeierstrassAffine(params)
    nb = 32 if name == "pallas" else 48
    pts = O.random_points_weierstrass(aff, n, seed=n)
    sc = O.random_scalars(n, params.q, seed=7 + n)
    want = O.msm(aff, sc, pts)
    port = Port(name)
    prepared = port.prepare_points(I.points_le(pts, nb), n, threads)
    for form in (0, 1):
        x, y, z, _ = port.msm(I.scalars_le(sc), prepared, n, threads, c, form)
        assert not z and (x, y) == want


def test_port_kat_and_edge_cases():
    # scripts/zprize23/submission-test-bls377.ts:18-25 and the safe-addition cases
    params = O.BLS12_377
    aff = O.WeierstrassAffine(params)
    P = (111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740,
         76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740)
    port = Port("bls12-377")
    prep = port.prepare_points(I.points_le([P, P], 48), 2)
    x, y, z, _ = port.msm(I.scalars_le([2, params.q - 1]), prep, 2, 1, 3)
    assert (x, y) == P and not z
    prep = port.prepare_points(I.points_le([P] * 16, 48), 16)
    x, y, z, _ = port.msm(I.scalars_le([5] * 16), prep, 16, 2, 4)
    assert (x, y) == aff.scale(80, P)
    prep = port.prepare_points(I.points_le([P, aff.negate(P)], 48), 2)
    x, y, z, _ = port.msm(I.scalars_le([7, 7]), prep, 2, 1, 4)
    assert z and (x, y) == (0, 0)


@pytest.mark.parametrize("n,threads,c", [(1, 1, 0), (2, 1, 3), (40, 3, 5), (300, 8, 7)])
def test_port_twisted_edwards(n, threads, c):
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    pts = O.random_points_te(te, n, seed=n)
    sc = O.random_scalars(n, te.q, seed=9 + n)
    want = te.to_affine(O.msm(te, sc, [te.from_affine(p) for p in pts]))
    port = Port("ed-on-bls12-377")
    prep = port.prepare_points(I.points_le(pts, 32), n, threads)
    x, y, z, _ = port.msm(I.scalars_le(sc), prep, n, threads, c)
    assert (x, y) == want


def test_port_window_policy_matches_reference_table():
    # SURVEY appendix B
    assert Port("bls12-377").default_window(1 << 18) == 14
    assert Port("bls12-377").default_window(1 << 20) == 18
    assert Port("pallas").default_window(1 << 16) == 12
    assert Port("ed-on-bls12-377").default_window(1 << 16) == 14


M64 = 2 ** 64 - 1


def _splitmix(st):
    st = (st + 0x9E3779B97F4A7C15) & M64
    z = st
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return st, z ^ (z >> 31)


@pytest.mark.parametrize("name", ["bls12-377", "pallas", "bls12-381", "ed-on-bls12-377"])
def test_port_seeded_generators(name):
    """The port's input generators (what `bench.py --impl reference` feeds the CPU arm with) restate the CUDA
    generators' construction: point i = sum_k (w_k + 1) * h_k * G with 13-bit w_k from SplitMix64 -- checked here
    against scalar multiplication in the python oracle; scalars uniform below q; ranges of a larger set agree."""
    port = Port(name)
    seed, n, nb = 0xB212, 300, port.nbytes
    pts = port.random_points(n, seed, 4)
    hs = []
    for k in range(4):
        _, h = _splitmix(seed ^ (0xA5A5A5A5 + k))
        hs.append(h | 1)
    for i in (0, 1, 255, 256, 299):  # both sides of the generator's batch boundary
        _, r = _splitmix(((seed ^ 0x5EED) + (i + 1) * 0xD1342543DE82EF95) & M64)
        s = sum(hs[k] * (((r >> (13 * k)) & 8191) + 1) for k in range(4))
        got = (int.from_bytes(pts[2 * nb * i:2 * nb * i + nb], "little"),
               int.from_bytes(pts[2 * nb * i + nb:2 * nb * (i + 1)], "little"))
        if name == "ed-on-bls12-377":
            te = O.TwistedEdwards(O.ED_ON_BLS12_377)
            want = te.to_affine(te.scale(s % te.q, te.one))
        else:
            aff = O.WeierstrassAffine({"bls12-377": O.BLS12_377, "pallas": O.PALLAS, "bls12-381": O.BLS12_381}[name])
            want = aff.scale(s % aff.q, aff.one)
        assert got == tuple(want)
    sc = port.random_scalars(n, 77, 3)
    vals = [int.from_bytes(sc[32 * i:32 * i + 32], "little") for i in range(n)]
    assert max(vals) < port.q and len(set(vals)) == n
    assert port.random_points(40, seed, 2, first=260) == pts[260 * 2 * nb:]
    assert port.random_scalars(40, 77, 2, first=260) == sc[260 * 32:]
