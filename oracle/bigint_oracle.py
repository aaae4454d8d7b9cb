"""CPU oracle for the MSM hot path -- TEST INFRASTRUCTURE ONLY.

This file is a plain-Python (arbitrary precision int) restatement of the reference's
BigInt oracle and of the boundary formats of its wasm MSM.  It is imported only by
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py`; the product path (msm_zprize_b200 + libmsm_b200.so) never touches it.

Parity status: PINNED.  `tests/test_oracle.py` checks this oracle against every stored
vector the reference holds for the path: the two ZPrize known-answer
tests (scripts/zprize23/submission-test-bls377.ts:6-45, submission-test.ts:5-21), the three
curve generators (on-curve + subgroup, src/concrete/*.params.ts), and the recomputed GLV
lattice constants.  The reference itself (TypeScript + generated
wasm) cannot run in this image (no node), so beyond those vectors parity rests on this
restatement following the cited lines.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

# ---------------------------------------------------------------------------------------
# util  (src/util.ts:163-167  log2 = ceil(log2 n) as bit length of n-1 ... see below)
# ---------------------------------------------------------------------------------------


def log2(n: int) -> int:
    """src/util.ts:163-167 -- number of bits needed to index n values: ceil(log2(n)).

    The reference implements it as `(n - 1).toString(2).length` for n > 1 and 0 for n<=1...
    which equals the bit length of n-1.
    """
    n = int(n)
    if n <= 1:
        return 0
    return (n - 1).bit_length()


def mod(x: int, p: int) -> int:
    """src/bigint/field-util.ts:8-11"""
    return x % p


def inverse(x: int, p: int) -> int:
    """src/bigint/field.ts:117-125 (egcd based); Python's pow(x,-1,p) is the same map."""
    x %= p
    if x == 0:
        raise ZeroDivisionError("cannot invert 0")
    return pow(x, -1, p)


# ---------------------------------------------------------------------------------------
# curve constants  (src/concrete/*.params.ts)
# ---------------------------------------------------------------------------------------


@dataclass(frozen=True)
class WeierstrassParams:
    label: str
    p: int  # base field modulus
    q: int  # scalar field modulus (subgroup order)
    h: int  # cofactor
    b: int  # y^2 = x^3 + b  (a = 0)
    gx: int
    gy: int
    lam: int  # endomorphism scalar: lam*(x,y) = (beta*x, y)
    beta: int


@dataclass(frozen=True)
class TwistedEdwardsParams:
    label: str
    p: int
    q: int
    h: int
    d: int  # -x^2 + y^2 = 1 + d x^2 y^2
    gx: int
    gy: int


# src/concrete/bls12-377.params.ts:11-45
BLS12_377 = WeierstrassParams(
    label="bls12-377",
    p=0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001,
    q=0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    h=0x170B5D44300000000000000000000000,
    b=1,
    gx=0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF,
    gy=0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6,
    lam=0x12AB655E9A2CA55660B44D1E5C37B00114885F32400000000000000000000000,
    beta=0x1AE3A4617C510EABC8756BA8F8C524EB8882A75CC9BC8E359064EE822FB5BFFD1E945779FFFFFFFFFFFFFFFFFFFFFFF,
)

# src/concrete/pasta.params.ts:10-46
_PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
_PALLAS_Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
_pallas_lambda = pow(5, (_PALLAS_Q - 1) // 3, _PALLAS_Q)  # pasta.params.ts:22
_pallas_beta2 = pow(5, (_PALLAS_P - 1) // 3, _PALLAS_P)  # pasta.params.ts:31
_pallas_beta = (_pallas_beta2 * _pallas_beta2) % _PALLAS_P  # pasta.params.ts:32
PALLAS = WeierstrassParams(
    label="pallas",
    p=_PALLAS_P,
    q=_PALLAS_Q,
    h=1,
    b=5,
    gx=1,
    gy=0x1B74B5A30A12937C53DFA9F06378EE548F655BD4333D477119CF7A23CAED2ABB,
    lam=_pallas_lambda,
    beta=_pallas_beta,
)

# src/concrete/bls12-381.params.ts:6-57 (lambda2 = z^2 - 1, beta2)
BLS12_381 = WeierstrassParams(
    label="bls12-381",
    p=0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB,
    q=0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    h=0x396C8C005555E1568C00AAAB0000AAAB,
    b=4,
    gx=0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    gy=0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
    lam=0xD201000000010000 ** 2 - 1,
    beta=0x1A0111EA397FE699EC02408663D4DE85AA0D857D89759AD4897D29650FB85F9B409427EB4F49FFFD8BFD00000000AAAC,
)

# src/concrete/ed-on-bls12-377.params.ts:5-31
ED_ON_BLS12_377 = TwistedEdwardsParams(
    label="ed-on-bls12-377",
    p=0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001,
    q=0x4AAD957A68B2955982D1347970DEC005293A3AFC43C8AFEB95AEE9AC33FD9FF,
    h=4,
    d=3021,
    gx=0x9F1B5A5BAF6ACF06FED91C9AE9EBFA06068DD2835790980894E2328F3EBCA05,
    gy=0x9A20DF36571AC3CD906B256080BA8454453C177AAF3131BB50A67BF1A806781,
)

# ---------------------------------------------------------------------------------------
# affine short Weierstrass, a = 0  (src/bigint/affine-weierstrass.ts:44-119)
# A point is (x, y) or None for the point at infinity (`isZero: true`).
# ---------------------------------------------------------------------------------------

AffinePoint = Optional[Tuple[int, int]]


class WeierstrassAffine:
    def __init__(self, params: WeierstrassParams):
        self.params = params
        self.p = params.p
        self.q = params.q
        self.zero: AffinePoint = None
        self.one: AffinePoint = (params.gx, params.gy)
        self.scalar_bits = log2(params.q)  # Curve.Scalar.sizeInBits

    def add(self, P1: AffinePoint, P2: AffinePoint) -> AffinePoint:
        """affine-weierstrass.ts:44-69 (complete: zero, doubling, inverse cases)."""
        if P1 is None:
            return P2
        if P2 is None:
            return P1
        p = self.p
        x1, y1 = P1
        x2, y2 = P2
        if (x1 - x2) % p == 0:
            if (y1 - y2) % p == 0:
                return self.double(P1)
            assert (y1 + y2) % p == 0, "unreachable"
            return None
        d = inverse(x2 - x1, p)
        m = (y2 - y1) * d % p
        x3 = (m * m - x1 - x2) % p
        y3 = (m * (x1 - x3) - y1) % p
        return (x3, y3)

    def double(self, P: AffinePoint) -> AffinePoint:
        """affine-weierstrass.ts:74-87."""
        if P is None:
            return None
        p = self.p
        x, y = P
        if y % p == 0:
            return None
        d = inverse(2 * y, p)
        m = 3 * x * x * d % p
        x2 = (m * m - 2 * x) % p
        y2 = (m * (x - x2) - y) % p
        return (x2, y2)

    def negate(self, P: AffinePoint) -> AffinePoint:
        """affine-weierstrass.ts:92-95."""
        if P is None:
            return None
        return (P[0], (-P[1]) % self.p)

    def scale(self, s: int, P: AffinePoint) -> AffinePoint:
        """affine-weierstrass.ts:113-121 (MSB-first double-and-add)."""
        Q: AffinePoint = None
        for i in range(s.bit_length() - 1, -1, -1):
            Q = self.double(Q)
            if (s >> i) & 1:
                Q = self.add(Q, P)
        return Q

    def is_on_curve(self, P: AffinePoint) -> bool:
        """affine-weierstrass.ts:134-137."""
        if P is None:
            return True
        x, y = P
        return (y * y - x * x * x - self.params.b) % self.p == 0

    def is_in_subgroup(self, P: AffinePoint) -> bool:
        """affine-weierstrass.ts:139-141."""
        return self.scale(self.q, P) is None

    def endo(self, P: AffinePoint) -> AffinePoint:
        """src/wasm/curve.ts:90-103 -- (beta*x, y) == lambda*P on the subgroup."""
        if P is None:
            return None
        return (P[0] * self.params.beta % self.p, P[1])


# ---------------------------------------------------------------------------------------
# homogeneous projective short Weierstrass (src/bigint/projective-weierstrass.ts:33-115)
# (X, Y, Z), zero <=> Z == 0.
# ---------------------------------------------------------------------------------------

ProjPoint = Tuple[int, int, int]


class WeierstrassProjective:
    def __init__(self, params: WeierstrassParams):
        self.params = params
        self.p = params.p
        self.q = params.q
        self.zero: ProjPoint = (0, 1, 0)
        self.one: ProjPoint = (params.gx, params.gy, 1)
        self.scalar_bits = log2(params.q)

    def from_affine(self, P: AffinePoint) -> ProjPoint:
        return self.zero if P is None else (P[0], P[1], 1)

    def to_affine(self, P: ProjPoint) -> AffinePoint:
        """projective-weierstrass.ts:201-209."""
        X, Y, Z = P
        if Z % self.p == 0:
            return None
        zi = inverse(Z, self.p)
        return (X * zi % self.p, Y * zi % self.p)

    def add(self, P1: ProjPoint, P2: ProjPoint) -> ProjPoint:
        """projective-weierstrass.ts:33-82 (add-1998-cmo-2 + zero/double/inverse)."""
        p = self.p
        X1, Y1, Z1 = P1
        X2, Y2, Z2 = P2
        if Z1 % p == 0:
            return P2
        if Z2 % p == 0:
            return P1
        Y1Z2 = Y1 * Z2 % p
        X1Z2 = X1 * Z2 % p
        Z1Z2 = Z1 * Z2 % p
        u = (Y2 * Z1 - Y1Z2) % p
        uu = u * u % p
        v = (X2 * Z1 - X1Z2) % p
        if v == 0:
            if u == 0:
                return self.double(P1)
            return self.zero
        vv = v * v % p
        vvv = v * vv % p
        R = vv * X1Z2 % p
        A = (uu * Z1Z2 - vvv - 2 * R) % p
        X3 = v * A % p
        Y3 = (u * (R - A) - vvv * Y1Z2) % p
        Z3 = vvv * Z1Z2 % p
        return (X3, Y3, Z3)

    def double(self, P: ProjPoint) -> ProjPoint:
        """projective-weierstrass.ts:87-121 (dbl-1998-cmo-2, a = 0)."""
        p = self.p
        X1, Y1, Z1 = P
        if Z1 % p == 0:
            return self.zero
        w = 3 * X1 * X1 % p
        s = Y1 * Z1 % p
        ss = s * s % p
        sss = s * ss
        R = Y1 * s % p
        B = X1 * R % p
        h = (w * w - 8 * B) % p
        X3 = 2 * h * s % p
        Y3 = (w * (4 * B - h) - 8 * R * R) % p
        Z3 = 8 * sss % p
        return (X3, Y3, Z3)

    def is_equal(self, P1: ProjPoint, P2: ProjPoint) -> bool:
        return self.to_affine(P1) == self.to_affine(P2)


# ---------------------------------------------------------------------------------------
# twisted Edwards a = -1, extended coordinates (src/bigint/twisted-edwards.ts:36-95)
# ---------------------------------------------------------------------------------------

ExtPoint = Tuple[int, int, int, int]


class TwistedEdwards:
    def __init__(self, params: TwistedEdwardsParams):
        self.params = params
        self.p = params.p
        self.q = params.q
        self.k = 2 * params.d % params.p
        self.zero: ExtPoint = (0, 1, 1, 0)
        self.one: ExtPoint = self.from_affine((params.gx, params.gy))
        self.scalar_bits = log2(params.q)

    def from_affine(self, P: Tuple[int, int]) -> ExtPoint:
        """twisted-edwards.ts:36-38."""
        x, y = P
        return (x, y, 1, x * y % self.p)

    def to_affine(self, P: ExtPoint) -> Tuple[int, int]:
        """twisted-edwards.ts:39-45."""
        X, Y, Z, _ = P
        assert Z % self.p != 0, "Not an affine point"
        zi = inverse(Z, self.p)
        return (X * zi % self.p, Y * zi % self.p)

    def add(self, P1: ExtPoint, P2: ExtPoint) -> ExtPoint:
        """twisted-edwards.ts:52-86 (add-2008-hwcd-3, k = 2d, strongly unified)."""
        p = self.p
        X1, Y1, Z1, T1 = P1
        X2, Y2, Z2, T2 = P2
        A = (Y1 - X1) * (Y2 - X2) % p
        B = (Y1 + X1) * (Y2 + X2) % p
        C = T1 * T2 % p * self.k % p
        D = 2 * Z1 * Z2 % p
        E = (B - A) % p
        F = (D - C) % p
        G = (D + C) % p
        H = (B + A) % p
        return (E * F % p, G * H % p, F * G % p, E * H % p)

    def double(self, P: ExtPoint) -> ExtPoint:
        """twisted-edwards.ts:93-95."""
        return self.add(P, P)

    def negate(self, P: ExtPoint) -> ExtPoint:
        """twisted-edwards.ts:100-102."""
        X, Y, Z, T = P
        return ((-X) % self.p, Y, Z, (-T) % self.p)

    def scale(self, s: int, P: ExtPoint) -> ExtPoint:
        Q = self.zero
        for i in range(s.bit_length() - 1, -1, -1):
            Q = self.double(Q)
            if (s >> i) & 1:
                Q = self.add(Q, P)
        return Q

    def is_on_curve(self, P: ExtPoint) -> bool:
        x, y = self.to_affine(P)
        p = self.p
        return (-x * x + y * y - 1 - self.params.d * x * x % p * y * y) % p == 0

    def is_zero(self, P: ExtPoint) -> bool:
        return self.to_affine(P) == (0, 1)


# ---------------------------------------------------------------------------------------
# the oracle MSM  (src/bigint/msm.ts:8-53 -- unsigned-window Pippenger, generic in the curve)
# ---------------------------------------------------------------------------------------


def msm(curve, scalars: Sequence[int], points: Sequence) -> object:
    """src/bigint/msm.ts:8-53.  `curve` needs .zero, .add, .double, .scalar_bits."""
    N = len(scalars)
    assert N == len(points), "matching length"
    b = curve.scalar_bits
    c = max(log2(N) - 1, 1)
    c_mask = (1 << c) - 1
    K = -(-b // c)
    L = 1 << c
    partition_sums = []
    for k in range(K):
        buckets = [curve.zero] * (L - 1)
        for i in range(N):
            l = (scalars[i] >> (k * c)) & c_mask
            if l == 0:
                continue
            buckets[l - 1] = curve.add(buckets[l - 1], points[i])
        running = curve.zero
        triangle = curve.zero
        for l in range(L - 2, -1, -1):
            running = curve.add(running, buckets[l])
            triangle = curve.add(triangle, running)
        partition_sums.append(triangle)
    result = partition_sums[K - 1]
    for k in range(K - 2, -1, -1):
        for _ in range(c):
            result = curve.double(result)
        result = curve.add(result, partition_sums[k])
    return result


def msm_naive(curve, scalars: Sequence[int], points: Sequence) -> object:
    """sum_i s_i * P_i by double-and-add; independent cross-check of `msm`."""
    acc = curve.zero
    for s, P in zip(scalars, points):
        acc = curve.add(acc, curve.scale(s, P))
    return acc


# ---------------------------------------------------------------------------------------
# boundary formats: w-bit limbs, Montgomery radix  (src/bigint/field-util.ts:18-42,
# src/wasm/memory-helpers.ts:84-101, src/field-msm.ts:165-185)
# ---------------------------------------------------------------------------------------


@dataclass(frozen=True)
class MontParams:
    p: int
    w: int
    n: int  # limbs
    K: int  # n*w
    R: int  # 2^K
    length_p: int
    n_packed_bytes: int

    @property
    def size_field(self) -> int:
        return 4 * self.n  # bytes; each limb is stored in a u32


def montgomery_params(p: int, w: int = 29, min_extra_bits: int = 2) -> MontParams:
    """src/bigint/field-util.ts:18-42."""
    assert w <= 32
    length_p = log2(p)
    min_k = length_p + min_extra_bits
    n = -(-min_k // w)
    K = n * w
    return MontParams(p=p, w=w, n=n, K=K, R=1 << K, length_p=length_p,
                      n_packed_bytes=-(-length_p // 8))


def to_limbs(x: int, w: int, n: int) -> List[int]:
    """src/wasm/memory-helpers.ts:84-90 / src/util.ts bigintToLimbs: little-endian w-bit limbs."""
    mask = (1 << w) - 1
    out = []
    for _ in range(n):
        out.append(x & mask)
        x >>= w
    assert x == 0, "value does not fit"
    return out


def from_limbs(limbs: Sequence[int], w: int) -> int:
    """src/wasm/memory-helpers.ts:92-101."""
    x = 0
    for i, l in enumerate(limbs):
        x |= int(l) << (w * i)
    return x


def to_montgomery(x: int, mp: MontParams) -> int:
    """src/field-msm.ts:179-181: multiply(x, x, R^2) -> x*R mod p (result may be in [0,2p))."""
    return x * mp.R % mp.p


def from_montgomery(x: int, mp: MontParams) -> int:
    """src/field-msm.ts:182-185: multiply(x, x, 1) then reduce -> canonical."""
    return x * inverse(mp.R, mp.p) % mp.p


# --- point layouts (src/curve-affine.ts:20-52,77; curve-projective.ts:18; curve-twisted-edwards.ts:30-31)


def affine_size(mp: MontParams) -> int:
    return 2 * mp.size_field + 4


def projective_size(mp: MontParams) -> int:
    return 3 * mp.size_field + 4


def te_size(mp: MontParams) -> int:
    return 4 * mp.size_field


def encode_affine_limb29(points: Sequence[AffinePoint], mp: MontParams, unreduced=None):
    """Affine Weierstrass points -> the reference's in-memory layout (x | y | u8 isNonZero | pad),
    Montgomery form with w-bit limbs in u32 words (src/curve-affine.ts:20-52,
    Affine.writeBigint :235-252).  `unreduced`: optional iterable of bools -- add p to the stored
    coordinate (allowed: wasm multiply outputs are only < 2p).
    Returns a list of u32 words (length = len(points) * (2n+1)).
    """
    words: List[int] = []
    for i, P in enumerate(points):
        if P is None:
            words += [0] * (2 * mp.n) + [0]
            continue
        x = to_montgomery(P[0], mp)
        y = to_montgomery(P[1], mp)
        if unreduced is not None and unreduced[i]:
            if x + mp.p < 2 * mp.p:
                x += mp.p
            if y + mp.p < 2 * mp.p:
                y += mp.p
        words += to_limbs(x, mp.w, mp.n) + to_limbs(y, mp.w, mp.n) + [1]
    return words


def encode_te_limb29(points: Sequence[Tuple[int, int]], mp: MontParams):
    """TE affine (x,y) -> extended (X,Y,Z=1,T=xy) Montgomery w-bit limbs
    (src/curve-twisted-edwards.ts:30-31, writeBigint; parallel.ts:209-232)."""
    words: List[int] = []
    for (x, y) in points:
        for v in (x, y, 1, x * y % mp.p):
            words += to_limbs(to_montgomery(v, mp), mp.w, mp.n)
    return words


def encode_scalars_limb29(scalars: Sequence[int], mp: MontParams):
    """Scalars are stored as plain (non-Montgomery) w-bit limbs (src/scalar-glv.ts:60-66)."""
    words: List[int] = []
    for s in scalars:
        words += to_limbs(s, mp.w, mp.n)
    return words


def le_bytes(x: int, nbytes: int) -> bytes:
    """src/wasm/field-helpers.ts:211-301 packed little-endian byte string."""
    return int(x).to_bytes(nbytes, "little")


# ---------------------------------------------------------------------------------------
# GLV  (src/glv/glv.ts:21-50, src/wasm/glv.ts:35-63,187-226, src/glv/glv-test.ts:96-100,143-149)
# ---------------------------------------------------------------------------------------


def egcd_stop_early(l: int, p: int):
    """src/glv/glv.ts:21-50.  JS BigInt `/` truncates toward zero; all operands of `/` here are
    positive so Python's // agrees."""
    assert l <= p
    r0, r1 = p, l
    s0, s1 = 1, 0
    t0, t1 = 0, 1
    while r1 * r1 > p:
        quotient = r0 // r1
        r0, r1 = r1, r0 - quotient * r1
        s0, s1 = s1, s0 - quotient * s1
        t0, t1 = t1, t0 - quotient * t1
    quotient = r0 // r1
    r2 = r0 - quotient * r1
    t2 = t0 - quotient * t1
    v00, v10 = r1, -t1
    if max(r0, abs(t0)) <= max(r2, abs(t2)):
        v01, v11 = r0, -t0
    else:
        v01, v11 = r2, -t2
    return ((v00, v01), (v10, v11))


def _trunc_div(a: int, b: int) -> int:
    """JS BigInt division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


@dataclass(frozen=True)
class GlvParams:
    q: int
    lam: int
    w: int
    n: int
    n0: int
    m: int
    k: int
    v00: int
    v01: int
    v10: int
    v11: int
    det: int
    m0: int
    m1: int
    max_bits: int


def glv_params(q: int, lam: int, w: int = 29) -> GlvParams:
    """src/wasm/glv.ts:35-63 (constants) and :216-226 (maxBits); scalar limb count from
    src/scalar-glv.ts:36 (montgomeryParams(q, w, minExtraBits = 1))."""
    mp = montgomery_params(q, w, 1)
    n = mp.n
    n0 = -(-n // 2)
    m = n0 * w
    k = (n - n0) * w
    assert k <= m
    (v00, v01), (v10, v11) = egcd_stop_early(lam, q)
    det = v00 * v11 - v10 * v01
    m0 = _trunc_div((1 << (m + k)) * -v11, det)
    m1 = _trunc_div((1 << (m + k)) * v10, det)
    # upper bounds (src/wasm/glv.ts:216-226) -- evaluated in exact rationals here, the reference
    # uses doubles; the ceil(log2) outcome is what matters and is asserted in tests against the
    # values recorded in SURVEY appendix A.3 (126 / 127).
    from fractions import Fraction as Fr

    def js_rem(a, b):  # JS % takes the sign of the dividend
        return a - _trunc_div(a, b) * b

    m0_res = js_rem((1 << (m + k)) * -v11, det)
    m1_res = js_rem((1 << (m + k)) * v10, det)
    m0_err = abs(Fr(m0_res, det))
    m1_err = abs(Fr(m1_res, det))
    x0_err = Fr(1, 2) + Fr(m0, 1 << m) + m0_err * Fr(q, 1 << (m + k))
    x1_err = Fr(1, 2) + Fr(m1, 1 << m) + m1_err * Fr(q, 1 << (m + k))
    # note: the reference adds the *signed* m_i/2^m (negative for both curves) -- kept as is.
    max_s0 = x0_err * abs(v00) + x1_err * abs(v01)
    max_s1 = x0_err * abs(v10) + x1_err * abs(v11)
    max_bits = max(log2(int(abs(max_s0)) + 1), log2(int(abs(max_s1)) + 1))
    return GlvParams(q=q, lam=lam, w=w, n=n, n0=n0, m=m, k=k, v00=v00, v01=v01, v10=v10,
                     v11=v11, det=det, m0=m0, m1=m1, max_bits=max_bits)


def _sign(x: int) -> int:
    return -1 if x < 0 else 1


def _div_pow2_round(x: int, m: int) -> int:
    """src/glv/glv-test.ts:143-149 == wasm multiplyMsb rounding (src/wasm/glv.ts:187-214)."""
    round_up = (x >> (m - 1)) & 1
    return (x >> m) + round_up


def glv_decompose(s: int, g: GlvParams) -> Tuple[int, int]:
    """Signed (s0, s1) with s0 + s1*lambda == s (mod q).
    src/wasm/glv.ts:68-169 restated as in src/glv/glv-test.ts:96-100 (x_i takes the sign of m_i:
    glv.ts:103-104)."""
    x0 = _sign(g.m0) * _div_pow2_round(abs(g.m0) * (s >> g.k), g.m)
    x1 = _sign(g.m1) * _div_pow2_round(abs(g.m1) * (s >> g.k), g.m)
    s0 = g.v00 * x0 + g.v01 * x1 + s
    s1 = g.v10 * x0 + g.v11 * x1
    return s0, s1


# ---------------------------------------------------------------------------------------
# signed-digit slicing and window policy (src/msm-batched-affine.ts:91-98,172-200;
# src/msm-basic.ts:56-95; src/msm-common.ts:8-57)
# ---------------------------------------------------------------------------------------

_WINDOW_TABLE = {  # src/msm-common.ts:33-57
    "large": {},
    "large-affine": {14: 13, 15: 14, 16: 14, 17: 14, 18: 14, 19: 18, 20: 18},
    "small": {16: 14},
    "small-affine": {16: 12},
}


def window_size(field_bits: int, n: int) -> int:
    """src/msm-common.ts:8-13."""
    t = _WINDOW_TABLE["large" if field_bits > 260 else "small"]
    return t.get(n, max(n - 1, 1))


def window_size_affine(field_bits: int, n: int) -> int:
    """src/msm-common.ts:15-21."""
    t = _WINDOW_TABLE["large-affine" if field_bits > 260 else "small-affine"]
    return t.get(n, max(n - 1, 1))


def signed_digits(s: int, c: int, K: int) -> List[Tuple[int, int]]:
    """src/msm-batched-affine.ts:178-191 (same rule in msm-basic.ts:83-92).
    Returns [(l, carry)] per window: the point goes to bucket l (1..L) negated iff carry == 1;
    l == 0 means skipped."""
    L = 1 << (c - 1)
    out = []
    carry = 0
    mask = (1 << c) - 1
    for k in range(K):
        l = ((s >> (k * c)) & mask) + carry
        if l > L:
            l = 2 * L - l
            carry = 1
        else:
            carry = 0
        out.append((l, carry))
    assert carry == 0, "K = ceil((b+1)/c) guarantees no final carry"
    return out


def msm_glv_signed(aff: WeierstrassAffine, g: GlvParams, scalars: Sequence[int],
                   points: Sequence[AffinePoint], c: Optional[int] = None) -> AffinePoint:
    """Algorithm restatement of src/msm-batched-affine.ts:74-328 at the group-law level (GLV split,
    sign folding A.4, signed digits A.5, bucket sums, running-sum reduction A.7, Horner).
    Same result as `msm` -- used to cross-check intermediate stages (digits / bucket contents)."""
    N = len(scalars)
    n = log2(N)
    if c is None:
        c = window_size_affine(log2(aff.p), n)
    b = g.max_bits
    K = -(-(b + 1) // c)
    L = 1 << (c - 1)
    halves = []  # (|s|, point-with-sign-folded)
    for s, P in zip(scalars, points):
        s0, s1 = glv_decompose(s, g)
        P0 = P if s0 >= 0 else aff.negate(P)
        E = aff.endo(P)
        P2 = E if s1 >= 0 else aff.negate(E)
        halves.append((abs(s0), P0))
        halves.append((abs(s1), P2))
    buckets = [[None] * (L + 1) for _ in range(K)]
    for hs, P in halves:
        for k, (l, carry) in enumerate(signed_digits(hs, c, K)):
            if l == 0:
                continue
            Q = aff.negate(P) if carry else P
            buckets[k][l] = aff.add(buckets[k][l], Q)
    partial = []
    for k in range(K):
        row = None
        tri = None
        for l in range(L, 0, -1):
            row = aff.add(row, buckets[k][l])
            tri = aff.add(tri, row)
        partial.append(tri)
    res = partial[K - 1]
    for k in range(K - 2, -1, -1):
        for _ in range(c):
            res = aff.double(res)
        res = aff.add(res, partial[k])
    return res


def msm_basic_signed(curve, scalars: Sequence[int], points: Sequence, c: int, negate: Callable):
    """Algorithm restatement of src/msm-basic.ts:45-176 (no GLV; signed digits; running sums)."""
    b = curve.scalar_bits
    K = -(-(b + 1) // c)
    L = 1 << (c - 1)
    buckets = [[curve.zero] * (L + 1) for _ in range(K)]
    for s, P in zip(scalars, points):
        for k, (l, carry) in enumerate(signed_digits(s, c, K)):
            if l == 0:
                continue
            buckets[k][l] = curve.add(buckets[k][l], negate(P) if carry else P)
    partial = []
    for k in range(K):
        row = curve.zero
        tri = curve.zero
        for l in range(L, 0, -1):
            row = curve.add(row, buckets[k][l])
            tri = curve.add(tri, row)
        partial.append(tri)
    res = partial[K - 1]
    for k in range(K - 2, -1, -1):
        for _ in range(c):
            res = curve.double(res)
        res = curve.add(res, partial[k])
    return res


# ---------------------------------------------------------------------------------------
# deterministic synthetic inputs (the reference has no seeds: src/util.ts:226-233)
# ---------------------------------------------------------------------------------------


class SplitMix64:
    """Counter-based PRNG shared (bit for bit) with the C port and the CUDA generators."""

    def __init__(self, seed: int):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self) -> int:
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)


def random_scalars(n: int, q: int, seed: int) -> List[int]:
    """Distribution of src/curve-random.ts:151-194: 32 random bytes, top byte masked to the bit
    length of q, rejection-sampled until < q."""
    bits = log2(q)
    rng = SplitMix64(seed)
    out = []
    while len(out) < n:
        x = 0
        for j in range(4):
            x |= rng.next() << (64 * j)
        x &= (1 << bits) - 1
        if x < q:
            out.append(x)
    return out


def random_points_weierstrass(aff: WeierstrassAffine, n: int, seed: int) -> List[AffinePoint]:
    """Cheap deterministic subgroup points for small tests: P_i = r_i * G (r_i from SplitMix64).
    (Distribution differs from randomPointsFast, src/curve-random.ts:24-91, which sums table
    multiples of 5 random bases; for an MSM any set of distinct subgroup points is equivalent.)"""
    rng = SplitMix64(seed ^ 0xC0FFEE)
    G = aff.one
    # walk: P_0 = r*G, P_{i+1} = P_i + (small random multiple table)  -> cheap and distinct
    table = [None]
    for _ in range(15):
        table.append(aff.add(table[-1], G))
    base = aff.scale((rng.next() << 64 | rng.next()) % aff.q, G)
    steps = [aff.scale((rng.next() % aff.q) | 1, G) for _ in range(16)]
    out = []
    cur = base
    for _ in range(n):
        out.append(cur)
        cur = aff.add(cur, steps[rng.next() & 15])
    return out


def random_points_te(te: TwistedEdwards, n: int, seed: int) -> List[Tuple[int, int]]:
    rng = SplitMix64(seed ^ 0xEDED)
    G = te.one
    base = te.scale((rng.next() << 64 | rng.next()) % te.q, G)
    steps = [te.scale((rng.next() % te.q) | 1, G) for _ in range(16)]
    out = []
    cur = base
    for _ in range(n):
        out.append(te.to_affine(cur))
        cur = te.add(cur, steps[rng.next() & 15])
    return out
