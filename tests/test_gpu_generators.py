"""Device input generators (replacing randomPointsFast / randomScalars, src/curve-random.ts) and
size-independent properties at sizes the python oracle cannot reach."""
import numpy as np
import pytest

from oracle import bigint_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mz():
    import msm_zprize_b200 as m
    return m


def _gen(eng, mz, n, seed):
    pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
    d_pts = eng.dev_alloc(n * pb)
    d_sc = eng.dev_alloc(n * 32)
    eng.random_points_device(d_pts, n, seed)
    eng.random_scalars_device(d_sc, n, seed + 1)
    pts = eng.d2h(d_pts, n * pb).tobytes()
    sc = eng.d2h(d_sc, n * 32).tobytes()
    return d_pts, d_sc, pts, sc


@pytest.mark.parametrize("name,params", [("bls12-377", O.BLS12_377), ("pallas", O.PALLAS), ("bls12-381", O.BLS12_381)])
def test_generated_weierstrass_inputs(mz, name, params):
    # src/curve-projective.test.ts / curve-twisted-edwards.test.ts:162-198: on curve, in subgroup
    aff = O.WeierstrassAffine(params)
    n, nb = 200, (32 if name == "pallas" else 48)
    with mz.MsmEngine(name) as eng:
        d_pts, d_sc, pts, sc = _gen(eng, mz, n, 11)
        P = [(int.from_bytes(pts[i * 2 * nb:i * 2 * nb + nb], "little"),
              int.from_bytes(pts[i * 2 * nb + nb:(i + 1) * 2 * nb], "little")) for i in range(n)]
        S = [int.from_bytes(sc[i * 32:(i + 1) * 32], "little") for i in range(n)]
        assert all(aff.is_on_curve(p) for p in P)
        assert all(aff.is_in_subgroup(p) for p in P[:4])
        assert len(set(P)) == n
        assert all(s < params.q for s in S) and len(set(S)) == n
        eng.set_bases_device(d_pts, n)
        res = eng.run(d_sc, n, on_device=True)
        assert (res.x, res.y) == O.msm(aff, S, P)


def test_generated_te_inputs(mz):
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    n = 200
    with mz.MsmEngine("ed-on-bls12-377") as eng:
        d_pts, d_sc, pts, sc = _gen(eng, mz, n, 13)
        P = [(int.from_bytes(pts[i * 64:i * 64 + 32], "little"), int.from_bytes(pts[i * 64 + 32:(i + 1) * 64], "little"))
             for i in range(n)]
        S = [int.from_bytes(sc[i * 32:(i + 1) * 32], "little") for i in range(n)]
        assert all(te.is_on_curve(te.from_affine(p)) for p in P)
        assert te.is_zero(te.scale(te.q, te.from_affine(P[0])))
        assert all(s < te.q for s in S)
        eng.set_bases_device(d_pts, n)
        res = eng.run(d_sc, n, on_device=True)
        assert (res.x, res.y) == te.to_affine(O.msm(te, S, [te.from_affine(p) for p in P]))


@pytest.mark.parametrize("name", ["bls12-377", "pallas", "bls12-381", "ed-on-bls12-377"])
def test_device_generators_equal_the_cpu_port_generators(mz, name):
    """`bench.py --impl reference` never loads the CUDA library: it draws its inputs from the CPU port's
    generators.  They must be byte-identical to the device generators, also for a range of a larger set."""
    from oracle.port import Port
    port = Port(name)
    n = 3000
    with mz.MsmEngine(name) as eng:
        d_pts, d_sc, pts, sc = _gen(eng, mz, n, 0xB212)
        assert pts == port.random_points(n, 0xB212, 4)
        assert sc == port.random_scalars(n, 0xB213, 4)
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        eng.random_points_device(d_pts, 500, 0xB212, first=2500)
        eng.random_scalars_device(d_sc, 500, 0xB213, first=2500)
        assert eng.d2h(d_pts, 500 * pb).tobytes() == pts[2500 * pb:]
        assert eng.d2h(d_sc, 500 * 32).tobytes() == sc[2500 * 32:]


@pytest.mark.parametrize("name,lg", [("bls12-377", 16), ("pallas", 16), ("ed-on-bls12-377", 16), ("bls12-381", 15)])
def test_large_n_properties(mz, name, lg):
    """n = 2^16: (i) the result does not depend on the window size, (ii) linearity:
    MSM(s, G) == MSM(s[:h], G[:h]) + MSM(s[h:], G[h:]) via run_partial + combine (the multi-GPU path),
    (iii) Weierstrass: the batched-affine and the projective bucket method agree."""
    n = 1 << lg
    with mz.MsmEngine(name) as eng:
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        d_pts, d_sc, _, _ = _gen(eng, mz, n, 21)
        eng.set_bases_device(d_pts, n)
        a = eng.run(d_sc, n, on_device=True)
        b = eng.run(d_sc, n, on_device=True, window_bits=9)
        assert (a.x, a.y, a.is_zero) == (b.x, b.y, b.is_zero)
        if name != "ed-on-bls12-377":
            p = eng.run(d_sc, n, on_device=True, form=mz.FORM_PROJECTIVE)
            assert (a.x, a.y) == (p.x, p.y)
        # two halves, combined like two GPUs would
        h = n // 2 + 123
        pbytes = eng.partial_bytes()
        d_part = eng.dev_alloc(2 * pbytes)
        eng.run_partial(d_sc, h, d_part, on_device=True)
        eng.set_bases_device(d_pts + h * pb, n - h)
        eng.run_partial(d_sc + h * 32, n - h, d_part + pbytes, on_device=True)
        c = eng.combine(d_part, 2)
        assert (a.x, a.y, a.is_zero) == (c.x, c.y, c.is_zero)


@pytest.mark.parametrize("name,lg,c2", [("pallas", 20, 13), ("ed-on-bls12-377", 22, 16), ("bls12-377", 20, 14),
                                        ("bls12-377", 22, 18), ("bls12-377", 24, 19)])
def test_baseline_config_sizes_properties(mz, name, lg, c2):
    """BASELINE.json configs[2..4] at their full sizes (Pallas 2^20, ed-on-bls12-377 2^22, BLS12-377 2^20..2^24),
    where no CPU checker finishes in seconds: size-independent properties of the exact result --
    a different window size gives the same point, and the sum over two unequal ranges (run_partial + combine,
    the multi-GPU path) equals the whole."""
    n = 1 << lg
    with mz.MsmEngine(name) as eng:
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        d_pts = eng.dev_alloc(n * pb)
        d_sc = eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, n, 0xB200 + lg)
        eng.random_scalars_device(d_sc, n, 0x5CA1A + lg)
        eng.set_bases_device(d_pts, n)
        a = eng.run(d_sc, n, on_device=True)
        b = eng.run(d_sc, n, on_device=True, window_bits=c2)
        assert not a.is_zero
        assert (a.x, a.y, a.is_zero) == (b.x, b.y, b.is_zero)
        h = (n // 3) | 1
        pbytes = eng.partial_bytes()
        d_part = eng.dev_alloc(2 * pbytes)
        eng.run_partial(d_sc, h, d_part, on_device=True)
        eng.set_bases_device(d_pts + h * pb, n - h)
        eng.run_partial(d_sc + h * 32, n - h, d_part + pbytes, on_device=True)
        c = eng.combine(d_part, 2)
        assert (a.x, a.y, a.is_zero) == (c.x, c.y, c.is_zero)
        eng.dev_free(d_part)
        eng.dev_free(d_sc)
        eng.dev_free(d_pts)


@pytest.mark.parametrize("name,lg", [("bls12-377", 18), ("pallas", 18)])
def test_full_size_against_cpu_port(mz, name, lg):
    """BASELINE.json configs[1] size (BLS12-377 G1, n = 2^18): the GPU result must be bit-identical to the
    multi-threaded C++ port of the reference algorithm on the same inputs (the python oracle is too slow
    here; the port is pinned to it in tests/test_port.py)."""
    import os
    from oracle.port import Port
    n = 1 << lg
    with mz.MsmEngine(name) as eng:
        d_pts, d_sc, pts, sc = _gen(eng, mz, n, 0xB200 + lg)
        eng.set_bases_device(d_pts, n)
        got = eng.run(d_sc, n, on_device=True)
    port = Port(name)
    threads = os.cpu_count() or 1
    prep = port.prepare_points(pts, n, threads)
    x, y, z, _ = port.msm(sc, prep, n, threads)
    assert (got.x, got.y, got.is_zero) == (x, y, z)


@pytest.mark.parametrize("name,lg", [("pallas", 20), ("bls12-377", 20), ("ed-on-bls12-377", 20)])
def test_reference_ceiling_sizes_against_cpu_port(mz, name, lg):
    """The largest sizes the reference itself runs (doc/zprize23.md:27: 2^20; BASELINE.json configs 3-5): the GPU
    result must equal the C++ port of the reference algorithm on the same seeded inputs, with the reference's
    window table on the CPU side (BLS12-377 c=18, Pallas c=19, ed-on-bls12-377 c=19).  Inputs come from the two
    generators independently (device for the GPU, port for the CPU)."""
    import os
    from oracle.port import Port
    n = 1 << lg
    threads = os.cpu_count() or 1
    port = Port(name)
    pts = port.random_points(n, 0xB200 + lg, threads)
    sc = port.random_scalars(n, 0x5CA1A + lg, threads)
    x, y, z, _ = port.msm(sc, port.prepare_points(pts, n, threads), n, threads)
    with mz.MsmEngine(name) as eng:
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        d_pts = eng.dev_alloc(n * pb)
        d_sc = eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, n, 0xB200 + lg)
        eng.random_scalars_device(d_sc, n, 0x5CA1A + lg)
        eng.set_bases_device(d_pts, n)
        got = eng.run(d_sc, n, on_device=True)
        assert (got.x, got.y, got.is_zero) == (x, y, z)
        # and through the host-buffer entry point with the port's bytes (what a caller of compute_msm passes)
        got2 = eng.msm(sc, pts, n)
        assert (got2.x, got2.y, got2.is_zero) == (x, y, z)
        eng.dev_free(d_sc)
        eng.dev_free(d_pts)


@pytest.mark.parametrize("name", ["bls12-377", "pallas", "bls12-381", "ed-on-bls12-377"])
@pytest.mark.parametrize("lg", [12, 14])
def test_msm_test_ts_sizes_against_cpu_port(mz, name, lg):
    """src/msm.test.ts:35-83,94-119: N up to 2^12 (and 2^14) on all four curves, msmUnsafe == oracle and
    msmProjective == msmUnsafe -- here against the C++ port of the reference algorithm."""
    from oracle.port import Port
    n = 1 << lg
    with mz.MsmEngine(name) as eng:
        d_pts, d_sc, pts, sc = _gen(eng, mz, n, 100 + lg)
        eng.set_bases_device(d_pts, n)
        got = eng.run(d_sc, n, on_device=True)
        proj = eng.run(d_sc, n, on_device=True, form=mz.FORM_PROJECTIVE) if name != "ed-on-bls12-377" else got
    port = Port(name)
    prep = port.prepare_points(pts, n, 4)
    x, y, z, _ = port.msm(sc, prep, n, 4)
    assert (got.x, got.y, got.is_zero) == (x, y, z)
    assert (proj.x, proj.y) == (x, y)
