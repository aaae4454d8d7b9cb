"""The device math core (fp.cuh / ec.cuh / glv.cuh) compiled for the host, against the oracle.
CPU only; the shim library is built by __graft_entry__.build() (g++)."""
import ctypes
import os
import random

import pytest

from oracle import bigint_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "msm_zprize_b200", "csrc", "libmsm_b200_hostmath.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build_hostmath()
    return ctypes.CDLL(LIB)


FIELDS = {0: (O.BLS12_377.p, 12), 1: (O.PALLAS.p, 8), 2: (O.ED_ON_BLS12_377.p, 8), 3: (O.BLS12_381.p, 12)}
# internal Montgomery radix R = 2^(32 N)  (fp.cuh fe_mul)
RBITS = {0: 384, 1: 256, 2: 256, 3: 384}


def limbs(x, n):
    return (ctypes.c_uint32 * n)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def val(arr):
    return sum(int(v) << (32 * i) for i, v in enumerate(arr))


def _rbits(p):
    return 32 * -(-O.log2(p) // 32)


def mont(x, p, n):
    return x * (1 << _rbits(p)) % p


def unmont(x, p, n):
    return x * pow(1 << _rbits(p), -1, p) % p


def enc_fes(vals, p, n):
    words = []
    for v in vals:
        m = mont(v, p, n)
        words += [(m >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    return (ctypes.c_uint32 * len(words))(*words)


def fe_op(lib, field, op, a, b=0):
    p, n = FIELDS[field]
    out = (ctypes.c_uint32 * n)()
    assert lib.ht_fe_op(field, op, limbs(a, n), limbs(b, n), out) == 0
    return val(out)


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_field_sqr_carry_patterns(lib, field):
    """fe_sqr has its own cross-product / doubling / redc carry chains (fp.cuh): words of all ones,
    single set words and random values, against both the integers and fe_mul(a, a)."""
    p, n = FIELDS[field]
    Ri = pow(1 << RBITS[field], -1, p)
    rng = random.Random(0x5A5A + field)
    vals = []
    for i in range(n):
        vals.append((0xFFFFFFFF << (32 * i)) % p)
        vals.append(((1 << (32 * (i + 1))) - 1) % p)
        vals.append((p - 1) ^ (0xFFFFFFFF << (32 * i)) if ((p - 1) ^ (0xFFFFFFFF << (32 * i))) < p else p - 1 - i)
    for _ in range(3000):
        v = 0
        for i in range(n):
            v |= rng.choice([0, 0xFFFFFFFF, 0x80000000, 1, rng.getrandbits(32), rng.getrandbits(32)]) << (32 * i)
        vals.append(v % p)
    for a in vals:
        got = fe_op(lib, field, 6, a)
        assert got == a * a * Ri % p
        assert got == fe_op(lib, field, 0, a, a)


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_field_ops(lib, field):
    p, n = FIELDS[field]
    R = 1 << RBITS[field]
    Ri = pow(R, -1, p)
    rng = random.Random(field)
    edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, R % p, (1 << (32 * n - 32)) % p]
    vals = edge + [rng.randrange(p) for _ in range(60)]
    for a in vals:
        for b in vals[:12] + [rng.randrange(p) for _ in range(4)]:
            assert fe_op(lib, field, 0, a, b) == a * b * Ri % p
            assert fe_op(lib, field, 1, a, b) == (a + b) % p
            assert fe_op(lib, field, 2, a, b) == (a - b) % p
        assert fe_op(lib, field, 6, a) == a * a * Ri % p
        assert fe_op(lib, field, 7, a) == (-a) % p
        assert fe_op(lib, field, 4, a) == a * R % p
        assert fe_op(lib, field, 5, a) == a * Ri % p
    for a in vals[1:] + [rng.randrange(1, p) for _ in range(500)]:
        # Montgomery-domain inverse: (xR)^-1 * R^2  (safegcd), plain inverse, and the Fermat cross-check
        assert fe_op(lib, field, 3, a) == pow(a, -1, p) * R * R % p
        assert fe_op(lib, field, 9, a) == pow(a, -1, p)
    for a in vals[1:12]:
        assert fe_op(lib, field, 8, a) == pow(a, -1, p) * R * R % p


@pytest.mark.parametrize("field,params,b3", [(0, O.BLS12_377, 3), (1, O.PALLAS, 15), (3, O.BLS12_381, 12)])
def test_projective_complete_formulas(lib, field, params, b3):
    p, n = FIELDS[field]
    aff = O.WeierstrassAffine(params)
    rng = random.Random(5)
    pts = O.random_points_weierstrass(aff, 6, 3)

    def enc(P, z=None):
        if P is None:
            return enc_fes((0, 1, 0), p, n)
        z = z or rng.randrange(1, p)
        return enc_fes((P[0] * z % p, P[1] * z % p, z), p, n)

    def dec(out):
        X, Y, Z = (unmont(val(out[i * n:(i + 1) * n]), p, n) for i in range(3))
        if Z == 0:
            return None
        zi = pow(Z, -1, p)
        return (X * zi % p, Y * zi % p)

    cases = [(pts[0], pts[1]), (pts[2], pts[2]), (pts[3], aff.negate(pts[3])), (None, pts[4]),
             (pts[4], None), (None, None)]
    for P, Q in cases:
        out = (ctypes.c_uint32 * (3 * n))()
        assert lib.ht_proj_op(field, 0, enc(P), enc(Q), out) == 0
        assert dec(out) == aff.add(P, Q)
        if Q is not None:
            assert lib.ht_proj_op(field, 1, enc(P), enc(Q, 1), out) == 0
            assert dec(out) == aff.add(P, Q)
        assert lib.ht_proj_op(field, 2, enc(P), enc(P), out) == 0
        assert dec(out) == aff.double(P)
        if field == 0:
            # unreduced variants (BLS12-377: 7 spare bits), canonical operands and operands shifted by p (< 2p)
            def up(buf, mask):
                vals_ = [val(buf[i * n:(i + 1) * n]) for i in range(3)]
                vals_ = [v + p if (mask >> i) & 1 else v for i, v in enumerate(vals_)]
                words = []
                for v in vals_:
                    words += [(v >> (32 * k)) & 0xFFFFFFFF for k in range(n)]
                return (ctypes.c_uint32 * len(words))(*words)
            for mp, mq in [(0, 0), (7, 7), (5, 2), (2, 5)]:
                assert lib.ht_proj_op(field, 4, up(enc(P), mp), up(enc(Q), mq), out) == 0
                assert dec(out) == aff.add(P, Q)
                if Q is not None:
                    assert lib.ht_proj_op(field, 5, up(enc(P), mp), enc(Q, 1), out) == 0
                    assert dec(out) == aff.add(P, Q)
                assert lib.ht_proj_op(field, 6, up(enc(P), mp), enc(P), out) == 0
                assert dec(out) == aff.double(P)
        # to_affine (uses fe_inv)
        assert lib.ht_proj_op(field, 3, enc(P), enc(P), out) == 0
        if P is None:
            assert out[2 * n] == 0
        else:
            assert out[2 * n] == 1
            assert (unmont(val(out[0:n]), p, n), unmont(val(out[n:2 * n]), p, n)) == P


@pytest.mark.parametrize("field,params", [(0, O.BLS12_377), (1, O.PALLAS), (3, O.BLS12_381)])
def test_affine_add_cases(lib, field, params):
    p, n = FIELDS[field]
    aff = O.WeierstrassAffine(params)
    pts = O.random_points_weierstrass(aff, 5, 9)

    def enc(P):
        return enc_fes(P or (0, 0), p, n)

    cases = [(pts[0], pts[1]), (pts[1], pts[0]), (pts[2], pts[2]), (pts[3], aff.negate(pts[3])),
             (None, pts[4]), (pts[4], None), (None, None)]
    for P, Q in cases:
        out = (ctypes.c_uint32 * (2 * n))()
        flags = (1 if P is None else 0) | (2 if Q is None else 0)
        is_inf = lib.ht_aff_op(field, flags, enc(P), enc(Q), out)
        want = aff.add(P, Q)
        if want is None:
            assert is_inf == 1
        else:
            assert is_inf == 0
            assert (unmont(val(out[0:n]), p, n), unmont(val(out[n:2 * n]), p, n)) == want


def test_twisted_edwards_formulas(lib):
    p, n = FIELDS[2]
    te = O.TwistedEdwards(O.ED_ON_BLS12_377)
    rng = random.Random(1)
    aff_pts = O.random_points_te(te, 4, 2)

    def dec(out):
        return tuple(unmont(val(out[i * n:(i + 1) * n]), p, n) for i in range(4))

    def scaled(P):
        z = rng.randrange(1, p)
        return tuple(v * z % p for v in P)

    ext = [scaled(te.from_affine(a)) for a in aff_pts] + [te.zero]
    for P in ext:
        for Q in ext:
            out = (ctypes.c_uint32 * (4 * n))()
            assert lib.ht_ext_op(0, enc_fes(P, p, n), enc_fes(Q, p, n), out) == 0
            assert te.to_affine(dec(out)) == te.to_affine(te.add(P, Q))
        for a in aff_pts + [(0, 1)]:
            Q = te.from_affine(a)
            out = (ctypes.c_uint32 * (4 * n))()
            assert lib.ht_ext_op(1, enc_fes(P, p, n), enc_fes(Q, p, n), out) == 0
            got = dec(out)
            assert te.to_affine(got) == te.to_affine(te.add(P, Q))
            assert got[3] * got[2] % p == got[0] * got[1] % p  # T*Z == X*Y
            assert lib.ht_ext_op(2, enc_fes(P, p, n), enc_fes(Q, p, n), out) == 0
            assert te.to_affine(dec(out)) == te.to_affine(te.add(P, te.negate(Q)))


@pytest.mark.parametrize("curve,params", [(0, O.BLS12_377), (1, O.PALLAS), (3, O.BLS12_381)])
def test_glv_decompose_matches_oracle(lib, curve, params):
    # src/glv/glv-test.ts:83-125 -- device decomposition == BigInt restatement, sample by sample
    g = O.glv_params(params.q, params.lam)
    rng = random.Random(21)
    q = params.q
    edge = [0, 1, 2, q - 1, q - 2, params.lam, q // 2, (1 << 116) - 1, 1 << 116, (1 << 252) % q]
    for s in edge + [rng.randrange(q) for _ in range(20000)]:
        s0 = (ctypes.c_uint32 * 4)()
        s1 = (ctypes.c_uint32 * 4)()
        flags = lib.ht_glv(curve, limbs(s, 8), s0, s1)
        w0, w1 = O.glv_decompose(s, g)
        assert val(s0) == abs(w0) and val(s1) == abs(w1)
        assert (flags & 1) == (1 if w0 < 0 else 0) and (flags >> 1) == (1 if w1 < 0 else 0)
    # out-of-range input is reduced first
    s = q + 12345
    s0 = (ctypes.c_uint32 * 4)()
    s1 = (ctypes.c_uint32 * 4)()
    lib.ht_glv(curve, limbs(s, 8), s0, s1)
    w0, w1 = O.glv_decompose(12345, g)
    assert val(s0) == abs(w0) and val(s1) == abs(w1)
