"""compute_msm mirror (msm_zprize_b200/submission.py) -- the two submission self-tests of the reference
(scripts/zprize23/submission-test-bls377.ts:6-45, scripts/zprize23/submission-test.ts:5-21) in all
argument forms, and the host-side conversions."""
import random

import numpy as np
import pytest

from oracle import bigint_oracle as O
from tests import inputs as I

KAT_BLS_P = (
    111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740,
    76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740,
)
KAT_ED_P = (
    2796670805570508460920584878396618987767121022598342527208237783066948667246,
    8134280397689638111748378379571739274369602049665521098046934931245960532166,
)


def _u32(v, words):
    return np.array([(v >> (32 * i)) & 0xFFFFFFFF for i in range(words)], dtype=np.uint32)


def test_conversions_on_host():
    from msm_zprize_b200.submission import points_to_bytes, scalars_to_bytes
    import msm_zprize_b200 as mz
    sc = [0, 1, O.BLS12_377.q - 1, (1 << 256) - 1]
    assert scalars_to_bytes(sc) == I.scalars_le(sc)
    assert scalars_to_bytes([_u32(s, 8) for s in sc]) == I.scalars_le(sc)
    pts = [KAT_BLS_P, (3, 4)]
    as_big = [{"x": x, "y": y, "isZero": False} for x, y in pts]
    as_u32 = [{"x": _u32(x, 12), "y": _u32(y, 12)} for x, y in pts]
    assert points_to_bytes(as_big, 48) == (I.points_le(pts, 48), [0, 1])
    assert points_to_bytes(as_u32, 48) == (I.points_le(pts, 48), [0, 1])
    as_big[0]["isZero"] = True
    assert points_to_bytes(as_big, 48) == (I.points_le(pts[1:], 48), [1])
    with pytest.raises(mz.MsmError):
        scalars_to_bytes([-1])
    with pytest.raises(mz.MsmError):
        scalars_to_bytes([1 << 256])
    with pytest.raises(mz.MsmError):
        points_to_bytes([{"x": 1 << 384, "y": 0}], 48)
    with pytest.raises(mz.MsmError):
        points_to_bytes([{"x": _u32(1, 8), "y": _u32(1, 8)}], 48)


@pytest.mark.gpu
def test_submission_bls12_377_self_test():
    from msm_zprize_b200.submission import Submission
    q = O.BLS12_377.q
    aff = O.WeierstrassAffine(O.BLS12_377)
    point = {"x": KAT_BLS_P[0], "y": KAT_BLS_P[1], "isZero": False}
    with Submission("bls12-377") as sub:
        # 2*P + (-1)*P gives P again (:18-27)
        r = sub.compute_msm([point, point], [2, q - 1])
        assert (r["x"], r["y"]) == KAT_BLS_P
        # the same through the Buffer and the u32 forms
        r = sub.compute_msm(I.points_le([KAT_BLS_P] * 2, 48), I.scalars_le([2, q - 1]))
        assert (r["x"], r["y"]) == KAT_BLS_P
        r = sub.compute_msm([{"x": _u32(KAT_BLS_P[0], 12), "y": _u32(KAT_BLS_P[1], 12)}] * 2,
                            [_u32(2, 8), _u32(q - 1, 8)])
        assert (r["x"], r["y"]) == KAT_BLS_P
        # 1000 x the same point == scaling by the sum of scalars (:29-45)
        rng = random.Random(23)
        sc = [rng.randrange(q) for _ in range(1000)]
        r2 = sub.compute_msm([point] * 1000, sc)
        r3 = sub.compute_msm([point], [sum(sc) % q])
        assert (r2["x"], r2["y"]) == (r3["x"], r3["y"]) == aff.scale(sum(sc) % q, KAT_BLS_P)
        # zero points in the bigint form contribute nothing; an all-zero input is the zero point
        zero = {"x": 0, "y": 0, "isZero": True}
        r = sub.compute_msm([zero, point, zero], [5, 7, 9])
        assert (r["x"], r["y"]) == aff.scale(7, KAT_BLS_P)
        r = sub.compute_msm([zero], [5])
        assert r["isZero"]
        r = sub.compute_msm(b"", b"")
        assert r["isZero"]
        # random points (:37-39, commented out upstream because of its cost there)
        pts = O.random_points_weierstrass(aff, 300, seed=5)
        sc = O.random_scalars(300, q, seed=6)
        r = sub.compute_msm([{"x": x, "y": y, "isZero": False} for x, y in pts], sc)
        assert (r["x"], r["y"]) == O.msm(aff, sc, pts)


@pytest.mark.gpu
def test_submission_ed_on_bls12_377_self_test():
    from msm_zprize_b200.submission import Submission
    import msm_zprize_b200 as mz
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    with Submission("ed-on-bls12-377") as sub:
        r = sub.compute_msm(I.points_le([KAT_ED_P] * 2, 32), I.scalars_le([2, te.q - 1]))
        assert (r["x"], r["y"]) == KAT_ED_P
        rng = random.Random(24)
        sc = [rng.randrange(te.q) for _ in range(1000)]
        r2 = sub.compute_msm(I.points_le([KAT_ED_P] * 1000, 32), I.scalars_le(sc))
        r3 = sub.compute_msm(I.points_le([KAT_ED_P], 32), I.scalars_le([sum(sc) % te.q]))
        assert (r2["x"], r2["y"]) == (r3["x"], r3["y"]) == te.to_affine(te.scale(sum(sc) % te.q, te.from_affine(KAT_ED_P)))
        with pytest.raises(mz.MsmError):
            sub.compute_msm(I.points_le([KAT_ED_P], 32), I.scalars_le([1, 2]))
        with pytest.raises(mz.MsmError):
            sub.compute_msm(I.points_le([KAT_ED_P], 32), b"\x01" * 33)
