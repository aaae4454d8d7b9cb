"""GPU parity tests: the CUDA path through the C ABI against the oracle on the same inputs.
Bit-exact (integer work): normalised affine output must be identical
(src/msm.test.ts:65-82,115-118; scripts/msm-weierstrass.ts:97-107)."""
import random

import numpy as np
import pytest

from oracle import bigint_oracle as O
from tests import inputs as I

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mz():
    import msm_zprize_b200 as m
    return m


FIELDS = {0: (O.BLS12_377.p, 12), 1: (O.PALLAS.p, 8), 2: (O.ED_ON_BLS12_377.p, 8), 3: (O.BLS12_381.p, 12)}
# internal Montgomery radix R = 2^(32 N)  (fp.cuh fe_mul)
RBITS = {0: 384, 1: 256, 2: 256, 3: 384}


def _limbs(vals, n):
    return np.array([[(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)] for v in vals], dtype=np.uint32)


def _vals(arr):
    return [sum(int(w) << (32 * i) for i, w in enumerate(row)) for row in arr]


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_field_ops_on_device(mz, field):
    # src/field.test.ts:15-155 -- multiply / add / subtract / inverse against BigInt
    from msm_zprize_b200.engine import test_field_op
    p, n = FIELDS[field]
    R = 1 << RBITS[field]
    Ri = pow(R, -1, p)
    rng = random.Random(field)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p]
    a = edge + [rng.randrange(p) for _ in range(505)]
    b = [rng.randrange(p) for _ in range(len(a) - 7)] + edge
    A, B = _limbs(a, n), _limbs(b, n)
    assert _vals(test_field_op(0, field, 0, A, B)) == [x * y * Ri % p for x, y in zip(a, b)]
    assert _vals(test_field_op(0, field, 1, A, B)) == [(x + y) % p for x, y in zip(a, b)]
    assert _vals(test_field_op(0, field, 2, A, B)) == [(x - y) % p for x, y in zip(a, b)]
    # the squaring has its own carry chains (fp.cuh fe_sqr): words of all ones / single set words as well
    sq = list(a)
    for i in range(n):
        sq += [(0xFFFFFFFF << (32 * i)) % p, ((1 << (32 * (i + 1))) - 1) % p, (1 << (32 * i)) % p]
    for _ in range(400):
        v = 0
        for i in range(n):
            v |= rng.choice([0, 0xFFFFFFFF, 0x80000000, 1, rng.getrandbits(32)]) << (32 * i)
        sq.append(v % p)
    SQ = _limbs(sq, n)
    assert _vals(test_field_op(0, field, 4, SQ, SQ)) == [x * x * Ri % p for x in sq]
    nz = [x if x else 1 for x in a][:64]
    assert _vals(test_field_op(0, field, 3, _limbs(nz, n), _limbs(nz, n))) == [pow(x, -1, p) * R * R % p for x in nz]
    # the quad-cooperative inversion used at the top of the product tree (inv_quad.cuh); 200 values incl. edges,
    # a count that is not a multiple of the warp size
    nz2 = [x if x else 1 for x in a][:200] + [1, 2, p - 1, p - 2, (p - 1) // 2, R % p, 3]
    assert _vals(test_field_op(0, field, 5, _limbs(nz2, n), _limbs(nz2, n))) == [pow(x, -1, p) * R * R % p for x in nz2]


@pytest.mark.parametrize("name,params,c", [("bls12-377", O.BLS12_377, 13), ("pallas", O.PALLAS, 7),
                                           ("bls12-377", O.BLS12_377, 16)])
def test_glv_digits_on_device(mz, name, params, c):
    # src/glv/glv-test.ts:83-125 + src/msm-batched-affine.ts:178-191, digit by digit
    g = O.glv_params(params.q, params.lam)
    n = 300
    sc = [0, 1, params.q - 1, params.lam] + O.random_scalars(n - 4, params.q, seed=5)
    with mz.MsmEngine(name) as eng:
        dig = eng.test_digits(I.scalars_le(sc), n, c)
    K = -(-(g.max_bits + 1) // c)
    assert dig.shape == (2 * n, K)
    for i, s in enumerate(sc):
        halves = O.glv_decompose(s, g)
        for j, h in enumerate(halves):
            want = O.signed_digits(abs(h), c, K)
            for k, (l, carry) in enumerate(want):
                neg = (carry ^ (1 if h < 0 else 0)) if l else 0
                assert int(dig[2 * i + j, k]) == (l | (neg << 31)), (i, j, k)


def _check_weierstrass(mz, name, params, scalars, points, layout="le", form=None, c=0, unreduced=None):
    aff = O.WeierstrassAffine(params)
    n = len(scalars)
    nb = 32 if name == "pallas" else 48
    want = O.msm(aff, scalars, points) if n else None
    with mz.MsmEngine(name) as eng:
        if layout == "le":
            assert all(P is not None for P in points)
            res = eng.msm(I.scalars_le(scalars), I.points_le(points, nb), n, mz.LAYOUT_LE_BYTES, mz.LAYOUT_LE_BYTES,
                          form=form, window_bits=c)
        else:
            res = eng.msm(I.scalars_limb29(scalars, params.q), I.points_limb29(points, params.p, unreduced), n,
                          mz.LAYOUT_LIMB29_MONT, mz.LAYOUT_LIMB29_MONT, form=form, window_bits=c)
    if want is None:
        assert res.is_zero and res.x == 0 and res.y == 0
    else:
        assert not res.is_zero
        assert (res.x, res.y) == want
    return res


KAT_BLS_P = (
    111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740,
    76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740,
)
KAT_ED_P = (
    2796670805570508460920584878396618987767121022598342527208237783066948667246,
    8134280397689638111748378379571739274369602049665521098046934931245960532166,
)


def test_kat_bls12_377(mz):
    # scripts/zprize23/submission-test-bls377.ts:18-45
    q = O.BLS12_377.q
    res = _check_weierstrass(mz, "bls12-377", O.BLS12_377, [2, q - 1], [KAT_BLS_P, KAT_BLS_P])
    assert (res.x, res.y) == KAT_BLS_P
    rng = random.Random(1)
    sc = [rng.randrange(q) for _ in range(1000)]
    aff = O.WeierstrassAffine(O.BLS12_377)
    with mz.MsmEngine("bls12-377") as eng:
        r2 = eng.msm(I.scalars_le(sc), I.points_le([KAT_BLS_P] * 1000, 48), 1000)
    assert (r2.x, r2.y) == aff.scale(sum(sc) % q, KAT_BLS_P)


@pytest.mark.parametrize("name,params", [("bls12-377", O.BLS12_377), ("pallas", O.PALLAS), ("bls12-381", O.BLS12_381)])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 16, 64, 257, 1024])
def test_msm_weierstrass_sizes(mz, name, params, n):
    # src/msm.test.ts:35-83: N = 2^0 .. 2^12 (the oracle here is python: up to 2^10)
    aff = O.WeierstrassAffine(params)
    pts = O.random_points_weierstrass(aff, n, seed=n)
    sc = O.random_scalars(n, params.q, seed=1000 + n)
    _check_weierstrass(mz, name, params, sc, pts, "le")
    if n <= 64:
        _check_weierstrass(mz, name, params, sc, pts, "limb29", unreduced=[i % 2 == 1 for i in range(n)])
        _check_weierstrass(mz, name, params, sc, pts, "le", form=mz.FORM_PROJECTIVE)


@pytest.mark.parametrize("c", [1, 2, 5, 9, 14, 16])
def test_window_size_invariance(mz, c):
    # any c gives the same point (src/msm-batched-affine.ts:79-98 options.c)
    aff = O.WeierstrassAffine(O.BLS12_377)
    n = 40
    pts = O.random_points_weierstrass(aff, n, seed=77)
    sc = O.random_scalars(n, aff.q, seed=78)
    _check_weierstrass(mz, "bls12-377", O.BLS12_377, sc, pts, "le", c=c)
    _check_weierstrass(mz, "bls12-377", O.BLS12_377, sc, pts, "le", c=c, form=mz.FORM_PROJECTIVE)


def test_edge_cases_safe_additions(mz):
    # batchAddNew semantics (src/curve-affine.ts:376-458): doubling, P + (-P), zero points,
    # scalars 0 / 1 / q-1; src/bigint/msm.test.ts:35-57 identities
    params = O.BLS12_377
    aff = O.WeierstrassAffine(params)
    q = params.q
    pts = O.random_points_weierstrass(aff, 8, seed=3)
    P, Q = pts[0], pts[1]
    # all-equal points with equal scalars: every pair is a doubling
    _check_weierstrass(mz, "bls12-377", params, [5] * 16, [P] * 16, c=4)
    # P and -P with the same scalar: cancels to zero
    _check_weierstrass(mz, "bls12-377", params, [7, 7], [P, aff.negate(P)], c=4)
    _check_weierstrass(mz, "bls12-377", params, [7, 7, 9], [P, aff.negate(P), Q], c=4)
    # zero / one / q-1 scalars
    _check_weierstrass(mz, "bls12-377", params, [0, 0, 0], pts[:3])
    _check_weierstrass(mz, "bls12-377", params, [0, 1, q - 1, 2], [P, Q, Q, P])
    # zero points in the reference's in-memory layout (isNonZero flag = 0)
    _check_weierstrass(mz, "bls12-377", params, [3, 4, 5, 6], [P, None, Q, None], layout="limb29")
    _check_weierstrass(mz, "bls12-377", params, [3, 4], [None, None], layout="limb29")
    # empty input
    _check_weierstrass(mz, "bls12-377", params, [], [])


def _check_te(mz, scalars, points, layout="le", c=0):
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    n = len(scalars)
    want = te.to_affine(O.msm(te, scalars, [te.from_affine(P) for P in points])) if n else (0, 1)
    with mz.MsmEngine("ed-on-bls12-377") as eng:
        if layout == "le":
            res = eng.msm(I.scalars_le(scalars), I.points_le(points, 32), n, window_bits=c)
        else:
            res = eng.msm(I.scalars_limb29(scalars, te.q), I.te_points_limb29(points, te.p), n,
                          mz.LAYOUT_LIMB29_MONT, mz.LAYOUT_LIMB29_MONT, window_bits=c)
    assert (res.x, res.y) == want
    assert res.is_zero == (want == (0, 1))
    return res


def test_kat_ed_on_bls12_377(mz):
    # scripts/zprize23/submission-test.ts:13-21
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    res = _check_te(mz, [2, te.q - 1], [KAT_ED_P, KAT_ED_P])
    assert (res.x, res.y) == KAT_ED_P


@pytest.mark.parametrize("n", [1, 2, 4, 16, 64, 300, 1024])
def test_msm_twisted_edwards_sizes(mz, n):
    # src/msm.test.ts:94-119
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    pts = O.random_points_te(te, n, seed=n)
    sc = O.random_scalars(n, te.q, seed=2000 + n)
    _check_te(mz, sc, pts, "le")
    if n <= 64:
        _check_te(mz, sc, pts, "limb29", c=5)


def test_te_edge_cases(mz):
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    pts = O.random_points_te(te, 4, seed=9)
    P = pts[0]
    negP = te.to_affine(te.negate(te.from_affine(P)))
    _check_te(mz, [5] * 8, [P] * 8, c=3)
    _check_te(mz, [7, 7], [P, negP], c=3)
    _check_te(mz, [0, 0], pts[:2])
    _check_te(mz, [], [])


def test_error_behaviour(mz):
    # errors are codes + messages, never a crash (the reference throws: src/util.ts:256)
    with mz.MsmEngine("bls12-377") as eng:
        with pytest.raises(mz.MsmError):
            eng.run(b"\0" * 32, 1)  # run before set_bases
        with pytest.raises(mz.MsmError):
            eng.msm(b"\0" * 32, b"\0" * 96, 1, form=mz.FORM_TE_EXTENDED)
        with pytest.raises(mz.MsmError):
            eng.msm(b"\0" * 32, b"\0" * 96, 1, window_bits=99)
    with pytest.raises(mz.MsmError):
        mz.MsmEngine("bls12-377", device=99)


def test_resident_bases_and_fresh_scalars(mz):
    # benchmark shape of scripts/msm-weierstrass.ts:12-51: bases fixed, new scalars per run
    params = O.BLS12_377
    aff = O.WeierstrassAffine(params)
    n = 128
    pts = O.random_points_weierstrass(aff, n, seed=42)
    with mz.MsmEngine("bls12-377") as eng:
        eng.set_bases(I.points_le(pts, 48), n)
        for it in range(3):
            sc = O.random_scalars(n, params.q, seed=500 + it)
            res = eng.run(I.scalars_le(sc), n)
            assert (res.x, res.y) == O.msm(aff, sc, pts)
        # prefix of the resident bases
        sc = O.random_scalars(50, params.q, seed=600)
        res = eng.run(I.scalars_le(sc), 50)
        assert (res.x, res.y) == O.msm(aff, sc, pts[:50])


def test_skewed_scalars_all_equal(mz):
    """Every point has the same scalar: each window has a single bucket holding all n entries
    (src/bigint/msm.test.ts:35-57 'same scalar => s * sum(P)').  Exercises the deep pairwise tree of the
    batched-affine path and the virtual-bucket split of the generic bucket method."""
    n = 3000
    for name, params in (("bls12-377", O.BLS12_377), ("pallas", O.PALLAS)):
        aff = O.WeierstrassAffine(params)
        pts = O.random_points_weierstrass(aff, n, seed=31)
        s = O.random_scalars(1, params.q, seed=32)[0]
        total = None
        for P in pts:
            total = aff.add(total, P)
        want = aff.scale(s, total)
        nb = 32 if name == "pallas" else 48
        with mz.MsmEngine(name) as eng:
            for form in (mz.FORM_AFFINE_GLV, mz.FORM_PROJECTIVE):
                r = eng.msm(I.scalars_le([s] * n), I.points_le(pts, nb), n, form=form)
                assert (r.x, r.y) == want, (name, form)
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    pts = O.random_points_te(te, n, seed=33)
    s = O.random_scalars(1, te.q, seed=34)[0]
    total = te.zero
    for P in pts:
        total = te.add(total, te.from_affine(P))
    want = te.to_affine(te.scale(s, total))
    with mz.MsmEngine("ed-on-bls12-377") as eng:
        r = eng.msm(I.scalars_le([s] * n), I.points_le(pts, 32), n)
        assert (r.x, r.y) == want


def test_run_partial_without_sync_and_last_timing(mz):
    """msm_b200_run_partial(timing = NULL) returns in stream order without a host synchronisation; the combine
    queued behind it sees the partial, and msm_b200_last_timing() reports the phases of that MSM afterwards.
    ShardedMsm (world size 1 here) drives exactly that sequence."""
    from msm_zprize_b200.dist import ShardedMsm
    params = O.BLS12_377
    aff = O.WeierstrassAffine(params)
    n = 300
    pts = O.random_points_weierstrass(aff, n, seed=77)
    sc = O.random_scalars(n, params.q, seed=78)
    want = O.msm(aff, sc, pts)
    with mz.MsmEngine("bls12-377") as eng:
        eng.set_bases(I.points_le(pts, 48), n)
        part = eng.dev_alloc(eng.partial_bytes())
        assert eng.run_partial(I.scalars_le(sc), n, part, timing=False) is None
        res = eng.combine(part, 1)
        assert (res.x, res.y) == want
        tm = eng.last_timing()
        assert tm["n_windows"] > 0 and tm["window_bits"] > 0 and tm["kernel_launches"] > 0
        assert tm["accumulate_ms"] > 0 and tm["reduce_ms"] > 0
        # with a timing struct the same call synchronises and reports at once
        tm2 = eng.run_partial(I.scalars_le(sc), n, part)
        assert tm2["n_windows"] == tm["n_windows"] and tm2["accumulate_ms"] > 0 and tm2["h2d_ms"] >= 0
    sh = ShardedMsm("bls12-377", device=0)
    sh.set_bases(I.points_le(pts, 48), n)
    r = sh.msm(I.scalars_le(sc), n)
    assert (r.x, r.y) == want
    sh.engine.close()


def test_set_bases_async_then_run(mz):
    """msm_b200_set_bases_async: upload + ingest on the copy stream, consumed by the next run (the two-call
    form of what msm_b200_msm does); a second upload replaces the first."""
    params = O.PALLAS
    aff = O.WeierstrassAffine(params)
    n = 200
    pts = O.random_points_weierstrass(aff, n, seed=91)
    pts2 = O.random_points_weierstrass(aff, n, seed=92)
    sc = O.random_scalars(n, params.q, seed=93)
    with mz.MsmEngine("pallas") as eng:
        buf = I.points_le(pts, 32)
        eng.set_bases_async(buf, n)
        r = eng.run(I.scalars_le(sc), n)
        assert (r.x, r.y) == O.msm(aff, sc, pts)
        buf2 = I.points_le(pts2, 32)
        eng.set_bases_async(buf2, n)
        eng.set_bases_async(buf, n)  # never consumed upload is superseded cleanly
        eng.set_bases_async(buf2, n)
        r = eng.run(I.scalars_le(sc), n)
        assert (r.x, r.y) == O.msm(aff, sc, pts2)


def test_entry_count_limit_is_an_error_not_an_overflow(mz):
    """Sorted entries are addressed with 32 bits: a call whose digit count reaches 2^31 must fail loudly."""
    n = (1 << 23) + (1 << 17)  # 2n * 127 one-bit windows = 2.16e9 >= 2^31
    with mz.MsmEngine("bls12-377") as eng:
        d_pts = eng.dev_alloc(n * 96)
        d_sc = eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, 1 << 10, 1)   # contents do not matter: the check precedes any work
        eng.random_scalars_device(d_sc, 1 << 10, 2)
        eng.set_bases_device(d_pts, n)
        with pytest.raises(mz.MsmError) as e:
            eng.run(d_sc, n, on_device=True, window_bits=1)
        assert e.value.code == -1  # MSM_E_INVALID
        r = eng.run(d_sc, 1 << 10, on_device=True)             # the context stays usable
        assert not r.is_zero


def test_python_binding_refuses_short_buffers(mz):
    """A host buffer shorter than n * item bytes must be refused with MSM_E_INVALID before the copy could read
    past its end (the N-API addon bounds-checks its regions the same way)."""
    from msm_zprize_b200 import _lib as L
    with mz.MsmEngine("bls12-377") as eng:
        with pytest.raises(mz.MsmError) as e:
            eng.set_bases(bytes(96 * 3), 4)
        assert e.value.code == L.E_INVALID
        eng.set_bases(bytes(96 * 4), 4)
        with pytest.raises(mz.MsmError):
            eng.run(bytes(32 * 3), 4)
        with pytest.raises(mz.MsmError):
            eng.msm(bytes(32 * 4), bytes(96 * 3), 4)
        with pytest.raises(mz.MsmError):
            eng.set_bases(np.zeros(10, dtype=np.uint32), 4, mz.LAYOUT_LIMB29_MONT)
    buf = mz.PinnedBuffer(64)
    buf.free()
    assert buf.array is None


def test_registered_host_memory(mz):
    """msm_b200_host_register: memory the caller owns (the buffer behind a wasm memory) page-locked in place; the
    MSM from it equals the MSM from an ordinary buffer."""
    import ctypes as C
    from msm_zprize_b200 import _lib as L
    params = O.PALLAS
    aff = O.WeierstrassAffine(params)
    n = 64
    pts = O.random_points_weierstrass(aff, n, seed=12)
    sc = O.random_scalars(n, params.q, seed=13)
    buf = np.frombuffer(bytearray(I.scalars_le(sc)), dtype=np.uint8)
    L.check(L.lib().msm_b200_host_register(C.c_void_p(buf.ctypes.data), buf.nbytes))
    try:
        with mz.MsmEngine("pallas") as eng:
            eng.set_bases(I.points_le(pts, 32), n)
            r = eng.run(buf, n)
            assert (r.x, r.y) == O.msm(aff, sc, pts)
    finally:
        L.check(L.lib().msm_b200_host_unregister(C.c_void_p(buf.ctypes.data)))
    with pytest.raises(mz.MsmError):
        L.check(L.lib().msm_b200_host_unregister(C.c_void_p(buf.ctypes.data)))  # not registered any more: a code, no crash


def test_two_contexts_from_two_host_threads(mz):
    """Contexts are independent: two of them driven from two host threads at once (what the N-API addon's worker
    pool does with two curves) give the same points as one after the other."""
    import threading
    params = O.BLS12_377
    n = 1 << 14
    from oracle.port import Port
    port = Port("bls12-377")
    pts = port.random_points(n, 31, 4)
    scs = [port.random_scalars(n, 40 + i, 4) for i in range(6)]
    with mz.MsmEngine("bls12-377") as e0:
        e0.set_bases(pts, n)
        want = [(r.x, r.y) for r in (e0.run(s, n) for s in scs)]
    got = {}

    def work(tid, lo):
        with mz.MsmEngine("bls12-377") as e:
            e.set_bases(pts, n)
            for i in range(lo, lo + 3):
                r = e.run(scs[i], n) if i % 2 else e.msm(scs[i], pts, n)
                got[i] = (r.x, r.y)

    th = [threading.Thread(target=work, args=(t, 3 * t)) for t in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert [got[i] for i in range(6)] == want


def test_workspace_regrows_and_shrinks_across_calls(mz):
    """One context, many calls of very different sizes and layouts (buffers are grown on demand and reused)."""
    from oracle.port import Port
    port = Port("pallas")
    nmax = 1 << 15
    pts = port.random_points(nmax, 51, 4)
    sc = port.random_scalars(nmax, 52, 4)
    with mz.MsmEngine("pallas") as eng:
        for n in (100, nmax, 7, 1 << 14, 3000, nmax, 1, 1 << 13):
            eng.set_bases(pts, n)
            a = eng.run(sc, n)
            b = eng.run(sc, n, window_bits=7)
            assert (a.x, a.y, a.is_zero) == (b.x, b.y, b.is_zero), n
            assert (a.x, a.y, a.is_zero) == port.msm(sc[:32 * n], port.prepare_points(pts[:64 * n], n, 4), n, 4)[:3], n


def test_shared_bases_and_pipelined_msms(mz):
    """msm_b200_share_bases: a second context runs over the first one's resident bases (no copy), also at the same
    time from another host thread; the loan ends with the lender's next set_bases (MSM_E_STATE, no stale read)."""
    from oracle.port import Port
    from msm_zprize_b200 import _lib as L
    port = Port("bls12-377")
    n = (1 << 14) + 11
    pts = port.random_points(n, 61, 4)
    scs = [port.random_scalars(n, 70 + i, 4) for i in range(6)]
    prep = port.prepare_points(pts, n, 4)
    want = [port.msm(s, prep, n, 4)[:3] for s in scs]
    with mz.PipelinedMsm("bls12-377", depth=2) as pipe:
        pipe.set_bases(pts, n)
        got = pipe.map(scs, n)
        assert [(r.x, r.y, r.is_zero) for r in got] == want
        assert all(r.timing["shared_buckets"] == 1 for r in got)
        owner, borrower = pipe.engines
        pts2 = port.random_points(n, 62, 4)
        owner.set_bases(pts2, n)  # the lender replaces its bases: the borrower must refuse, not read the new set
        with pytest.raises(mz.MsmError) as e:
            borrower.run(scs[0], n)
        assert e.value.code == L.E_STATE
        borrower.share_bases(owner)
        r = borrower.run(scs[0], n)
        assert (r.x, r.y, r.is_zero) == port.msm(scs[0], port.prepare_points(pts2, n, 4), n, 4)[:3]
        borrower.set_bases(pts, n)  # bases of its own again
        r = borrower.run(scs[1], n)
        assert (r.x, r.y, r.is_zero) == want[1]
    with mz.MsmEngine("bls12-377") as a, mz.MsmEngine("pallas") as b:
        with pytest.raises(mz.MsmError):
            a.share_bases(b)  # different curve
    # the lender is destroyed first: the borrower is left without bases (a code, no dangling read)
    lender, borrower = mz.MsmEngine("bls12-377"), mz.MsmEngine("bls12-377")
    lender.set_bases(pts, n)
    borrower.share_bases(lender)
    lender.close()
    with pytest.raises(mz.MsmError) as e:
        borrower.run(scs[0], n)
    assert e.value.code == L.E_STATE
    borrower.set_bases(pts, n)
    r = borrower.run(scs[0], n)
    assert (r.x, r.y, r.is_zero) == want[0]
    borrower.close()
