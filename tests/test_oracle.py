"""Pins oracle/bigint_oracle.py against every stored vector the reference holds for the MSM path
(SURVEY.md section 8c).  CPU only."""
import random

import pytest

from oracle import bigint_oracle as O

# scripts/zprize23/submission-test-bls377.ts:6-10
KAT_BLS_P = (
    111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740,
    76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740,
)
# scripts/zprize23/submission-test.ts:5-10
KAT_ED_P = (
    2796670805570508460920584878396618987767121022598342527208237783066948667246,
    8134280397689638111748378379571739274369602049665521098046934931245960532166,
)
KAT_ED_T = 3446088593515175914550487355059397868296219355049460558182099906777968652023


def test_generators_on_curve_and_in_subgroup():
    for params in (O.BLS12_377, O.PALLAS, O.BLS12_381):
        aff = O.WeierstrassAffine(params)
        assert aff.is_on_curve(aff.one)
        assert aff.is_in_subgroup(aff.one)
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    assert te.is_on_curve(te.one)
    assert te.is_zero(te.scale(te.q, te.one))


def test_endomorphism_constants():
    # src/concrete/bls12-377.params.ts:47-70 (debug block) and pasta.params.ts:20-35
    for params in (O.BLS12_377, O.PALLAS, O.BLS12_381):
        aff = O.WeierstrassAffine(params)
        assert pow(params.lam, 3, params.q) == 1 and params.lam != 1
        assert pow(params.beta, 3, params.p) == 1 and params.beta != 1
        lamG = aff.scale(params.lam, aff.one)
        assert lamG == aff.endo(aff.one)


def test_kat_bls12_377_two_points():
    # submission-test-bls377.ts:18-25: 2*P + (q-1)*P == P
    aff = O.WeierstrassAffine(O.BLS12_377)
    P = KAT_BLS_P
    assert aff.is_on_curve(P) and aff.is_in_subgroup(P)
    res = O.msm(aff, [2, aff.q - 1], [P, P])
    assert res == P
    proj = O.WeierstrassProjective(O.BLS12_377)
    resp = O.msm(proj, [2, aff.q - 1], [proj.from_affine(P)] * 2)
    assert proj.to_affine(resp) == P


def test_kat_bls12_377_same_points():
    # submission-test-bls377.ts:28-45: 1000 x P with random s_i == (sum s_i) * P   (n reduced: 200)
    aff = O.WeierstrassAffine(O.BLS12_377)
    rng = random.Random(7)
    n = 200
    scalars = [rng.randrange(aff.q) for _ in range(n)]
    lhs = O.msm(aff, scalars, [KAT_BLS_P] * n)
    rhs = aff.scale(sum(scalars) % aff.q, KAT_BLS_P)
    assert lhs == rhs


def test_kat_ed_on_bls12_377():
    # submission-test.ts:13-21
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    P = te.from_affine(KAT_ED_P)
    assert P[3] == KAT_ED_T
    assert te.is_on_curve(P)
    res = O.msm(te, [2, te.q - 1], [P, P])
    assert te.to_affine(res) == KAT_ED_P


def test_glv_constants_match_survey_appendix():
    g = O.glv_params(O.BLS12_377.q, O.BLS12_377.lam)
    assert (g.n, g.n0, g.m, g.k) == (9, 5, 145, 116)
    assert g.v00 == 1
    assert g.v01 == 0x452217CC900000010A11800000000001
    assert g.v10 == -0x452217CC900000010A11800000000000
    assert g.v11 == 1
    assert g.det == O.BLS12_377.q
    assert g.m0 == -0x1B6
    assert g.m1 == -0x767EF552D3FA6E2C0FEE5DA655F20303CF
    assert g.max_bits == 126
    h = O.glv_params(O.PALLAS.q, O.PALLAS.lam)
    assert h.v00 == 0x49E69D1640F049157FCAE1C700000001
    assert h.v01 == 0x49E69D1640A899538CB1279300000000
    assert h.v10 == -h.v01
    assert h.v11 == 0x93CD3A2C8198E2690C7C095A00000001
    assert h.det == O.PALLAS.q
    assert h.m0 == -0x49E69D1640CC7134863E04AD0000000058
    assert h.m1 == -0x24F34E8B20544CA9C65893C97FFFFFFFEC
    assert h.max_bits == 127
    for gg in (g, h):
        assert (gg.v00 + gg.lam * gg.v10) % gg.q == 0
        assert (gg.v01 + gg.lam * gg.v11) % gg.q == 0


def test_glv_decomposition_valid_and_bounded():
    # src/glv/glv-test.ts:83-125
    rng = random.Random(11)
    for params in (O.BLS12_377, O.PALLAS):
        g = O.glv_params(params.q, params.lam)
        edge = [0, 1, 2, params.q - 1, params.q - 2, params.lam, params.q // 2]
        for s in edge + [rng.randrange(params.q) for _ in range(3000)]:
            s0, s1 = O.glv_decompose(s, g)
            assert (s0 + s1 * params.lam) % params.q == s
            assert abs(s0) < (1 << g.max_bits) and abs(s1) < (1 << g.max_bits)


def test_signed_digits_recompose():
    rng = random.Random(3)
    for c in (1, 2, 5, 13, 14, 16):
        b = 126
        K = -(-(b + 1) // c)
        L = 1 << (c - 1)
        for _ in range(200):
            s = rng.randrange(1 << b)
            total = 0
            for k, (l, carry) in enumerate(O.signed_digits(s, c, K)):
                assert 0 <= l <= L
                total += (-l if carry else l) << (k * c)
            # digit with carry==1 means "use -P" for bucket l; the borrowed 2^c moves to the next window
            assert total == s


def test_window_tables():
    # SURVEY appendix B
    assert O.window_size_affine(377, 16) == 14
    assert O.window_size_affine(377, 18) == 14
    assert O.window_size_affine(377, 20) == 18
    assert O.window_size_affine(377, 22) == 21
    assert O.window_size_affine(255, 16) == 12
    assert O.window_size_affine(255, 20) == 19
    assert O.window_size(253, 16) == 14
    assert O.window_size(253, 22) == 21


def test_montgomery_params():
    # SURVEY appendix A.1
    assert O.montgomery_params(O.BLS12_377.p).n == 14
    assert O.montgomery_params(O.PALLAS.p).n == 9
    assert O.montgomery_params(O.ED_ON_BLS12_377.p).n == 9
    for q in (O.BLS12_377.q, O.PALLAS.q, O.ED_ON_BLS12_377.q):
        assert O.montgomery_params(q, 29, 1).n == 9
    mp = O.montgomery_params(O.BLS12_377.p)
    assert O.affine_size(mp) == 116 and O.projective_size(mp) == 172
    assert O.te_size(O.montgomery_params(O.ED_ON_BLS12_377.p)) == 144


@pytest.mark.parametrize("n", [1, 2, 5, 16, 33])
def test_msm_variants_agree(n):
    # src/bigint/msm.test.ts:18-101 -- pippenger == naive; affine == projective; GLV/signed == plain
    for params in (O.BLS12_377, O.PALLAS):
        aff = O.WeierstrassAffine(params)
        proj = O.WeierstrassProjective(params)
        g = O.glv_params(params.q, params.lam)
        pts = O.random_points_weierstrass(aff, n, seed=n)
        sc = O.random_scalars(n, params.q, seed=100 + n)
        a = O.msm(aff, sc, pts)
        assert a == O.msm_naive(aff, sc, pts)
        assert a == proj.to_affine(O.msm(proj, sc, [proj.from_affine(P) for P in pts]))
        assert a == O.msm_glv_signed(aff, g, sc, pts, c=4)
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    pts = [te.from_affine(P) for P in O.random_points_te(te, n, seed=n)]
    sc = O.random_scalars(n, te.q, seed=200 + n)
    a = te.to_affine(O.msm(te, sc, pts))
    assert a == te.to_affine(O.msm_naive(te, sc, pts))
    assert a == te.to_affine(O.msm_basic_signed(te, sc, pts, 5, te.negate))


def test_limb29_roundtrip():
    mp = O.montgomery_params(O.BLS12_377.p)
    aff = O.WeierstrassAffine(O.BLS12_377)
    pts = O.random_points_weierstrass(aff, 3, seed=1) + [None]
    words = O.encode_affine_limb29(pts, mp, unreduced=[False, True, False, False])
    assert len(words) == 4 * (2 * mp.n + 1)
    for i, P in enumerate(pts):
        w = words[i * 29:(i + 1) * 29]
        if P is None:
            assert w[28] == 0
            continue
        x = O.from_montgomery(O.from_limbs(w[:14], 29), mp)
        y = O.from_montgomery(O.from_limbs(w[14:28], 29), mp)
        assert (x, y) == P and w[28] == 1
