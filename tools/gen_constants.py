#!/usr/bin/env python3
"""Generates msm_zprize_b200/csrc/constants.cuh (field / curve / GLV constants as 32-bit limbs).

Self-contained on purpose (does not import oracle/): the product's constants must not depend on
test infrastructure.  tests/test_constants.py re-derives every value through the oracle and
compares.  Curve parameters: /root/reference src/concrete/bls12-377.params.ts:11-45,
pasta.params.ts:10-46, ed-on-bls12-377.params.ts:5-31.  GLV lattice: src/glv/glv.ts:21-50,
src/wasm/glv.ts:45-48.
"""
import os

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "msm_zprize_b200", "csrc", "constants.cuh")

BLS377_P = 0x01AE3A4617C510EAC63B05C06CA1493B1A22D9F300F5138F1EF3622FBA094800170B5D44300000008508C00000000001
BLS377_R = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001
BLS377_LAMBDA = 0x12AB655E9A2CA55660B44D1E5C37B00114885F32400000000000000000000000
BLS377_BETA = 0x1AE3A4617C510EABC8756BA8F8C524EB8882A75CC9BC8E359064EE822FB5BFFD1E945779FFFFFFFFFFFFFFFFFFFFFFF
PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
PALLAS_LAMBDA = pow(5, (PALLAS_Q - 1) // 3, PALLAS_Q)
PALLAS_BETA = pow(pow(5, (PALLAS_P - 1) // 3, PALLAS_P), 2, PALLAS_P)
BLS381_P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
BLS381_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS381_LAMBDA = 0xD201000000010000 ** 2 - 1
BLS381_BETA = 0x1A0111EA397FE699EC02408663D4DE85AA0D857D89759AD4897D29650FB85F9B409427EB4F49FFFD8BFD00000000AAAC
BLS381_GX = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
BLS381_GY = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
ED_P = BLS377_R
ED_Q = 0x4AAD957A68B2955982D1347970DEC005293A3AFC43C8AFEB95AEE9AC33FD9FF
ED_D = 3021
BLS377_GX = 0x008848DEFE740A67C8FC6225BF87FF5485951E2CAA9D41BB188282C8BD37CB5CD5481512FFCD394EEAB9B16EB21BE9EF
BLS377_GY = 0x01914A69C5102EFF1F674F5D30AFEEC4BD7FB348CA3E52D96D182AD44FB82305C2FE3D3634A9591AFD82DE55559C8EA6
PALLAS_GX = 1
PALLAS_GY = 0x1B74B5A30A12937C53DFA9F06378EE548F655BD4333D477119CF7A23CAED2ABB
ED_GX = 0x9F1B5A5BAF6ACF06FED91C9AE9EBFA06068DD2835790980894E2328F3EBCA05
ED_GY = 0x9A20DF36571AC3CD906B256080BA8454453C177AAF3131BB50A67BF1A806781


def limbs(x, n):
    out = [(x >> (32 * i)) & 0xFFFFFFFF for i in range(n)]
    assert x >> (32 * n) == 0
    return out


def arr(name, vals):
    body = ", ".join("0x%08xu" % v for v in vals)
    return ("  MSM_HD static uint32_t %s(int i) {\n    constexpr uint32_t t[%d] = {%s};\n"
            "    return t[i];\n  }\n" % (name, len(vals), body))


def ceil_log2(n):
    return (n - 1).bit_length() if n > 1 else 0


def field_struct(name, p, n32, extra=None, w29_extra_bits=2):
    rbits = 32 * n32
    R = 1 << rbits
    m0 = (-pow(p, -1, 1 << 32)) % (1 << 32)
    n29 = -(-(ceil_log2(p) + w29_extra_bits) // 29)  # limbs of the reference's in-memory format
    pl = ", ".join("0x%08xu" % v for v in limbs(p, n32))
    # The Montgomery factor M0 = -1/p mod 2^32 is 0xffffffff for all three fields.  If ptxas can
    # see that, it turns m = t0 * M0 into a negation and then refuses to fuse the m*p products into
    # IMAD.WIDE (it emits IMAD + IMAD.HI, 3 issue slots instead of 1; measured 2.2x on the
    # Montgomery product).  So on the device M0 is read from global memory through a pure
    # (CSE-able) asm load, which keeps its value opaque; the modulus limbs stay immediates.
    s = "#ifdef __CUDACC__\nstatic __device__ uint32_t %s_M0_g[1] = {0x%08xu};\n#endif\n" % (name, m0)
    s += "struct %s {\n  static constexpr int N = %d;\n  static constexpr int N29 = %d;\n" % (name, n32, n29)
    s += "  static constexpr int BITS = %d;\n" % ceil_log2(p)
    s += "  static constexpr int RBITS = %d;\n" % rbits
    s += "  static constexpr uint32_t M0 = 0x%08xu;\n" % m0
    s += ("  MSM_HD static uint32_t M0v() {\n#ifdef __CUDA_ARCH__\n    uint32_t r;\n"
          "    asm(\"ld.global.nc.u32 %%0, [%%1];\" : \"=r\"(r) : \"l\"(%s_M0_g));\n    return r;\n#else\n"
          "    return M0;\n#endif\n  }\n" % name)
    s += arr("P", limbs(p, n32))
    s += arr("ONE", limbs(R % p, n32))
    s += arr("R2", limbs(R * R % p, n32))
    s += arr("R3", limbs(R * R * R % p, n32))
    s += arr("PM2", limbs(p - 2, n32))
    # Unreduced ("lazy") arithmetic in the one-warp tail: with SPARE = floor(2^(32 n) / p) >= 64 sums of a few
    # dozen p still fit the limbs and fe_mul reduces any a * b with (a/p)(b/p) < SPARE.  k*p for a - b + k*p.
    spare = R // p
    s += "  static constexpr bool LAZY = %s;  // floor(2^%d / p) = %d\n" % ("true" if spare >= 64 else "false", rbits, spare)
    for k in (2, 3, 9):
        s += arr("P%d" % k, limbs(k * p if k * p < R else 0, n32))
    # limb29 Montgomery (R29 = 2^(29 n29)) <-> limb32 Montgomery conversion multipliers
    s += arr("FROM29", limbs(pow(2, 2 * rbits - 29 * n29, p), n32))
    s += arr("TO29", limbs(pow(2, 29 * n29, p), n32))
    for k, v in (extra or {}).items():
        s += arr(k, limbs(v * R % p, n32))
    s += "};\n\n"
    return s


def egcd_stop_early(l, p):
    r0, r1, t0, t1 = p, l, 0, 1
    while r1 * r1 > p:
        qq = r0 // r1
        r0, r1 = r1, r0 - qq * r1
        t0, t1 = t1, t0 - qq * t1
    qq = r0 // r1
    r2, t2 = r0 - qq * r1, t0 - qq * t1
    v00, v10 = r1, -t1
    if max(r0, abs(t0)) <= max(r2, abs(t2)):
        v01, v11 = r0, -t0
    else:
        v01, v11 = r2, -t2
    return v00, v01, v10, v11


def tdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def glv_struct(name, q, lam):
    n = -(-(ceil_log2(q) + 1) // 29)
    n0 = -(-n // 2)
    m = n0 * 29
    k = (n - n0) * 29
    v00, v01, v10, v11 = egcd_stop_early(lam, q)
    det = v00 * v11 - v10 * v01
    m0 = tdiv((1 << (m + k)) * -v11, det)
    m1 = tdiv((1 << (m + k)) * v10, det)
    assert (v00 + lam * v10) % q == 0 and (v01 + lam * v11) % q == 0
    # upper bound on the half scalars (src/wasm/glv.ts:216-226), in exact rationals
    from fractions import Fraction as Fr

    def js_rem(a, b):
        return a - tdiv(a, b) * b
    m0e = abs(Fr(js_rem((1 << (m + k)) * -v11, det), det))
    m1e = abs(Fr(js_rem((1 << (m + k)) * v10, det), det))
    x0e = Fr(1, 2) + Fr(m0, 1 << m) + m0e * Fr(q, 1 << (m + k))
    x1e = Fr(1, 2) + Fr(m1, 1 << m) + m1e * Fr(q, 1 << (m + k))
    max_s0 = x0e * abs(v00) + x1e * abs(v01)
    max_s1 = x0e * abs(v10) + x1e * abs(v11)
    maxbits = max(ceil_log2(int(abs(max_s0)) + 1), ceil_log2(int(abs(max_s1)) + 1))
    assert maxbits <= 127
    s = "struct %s {\n" % name
    s += "  static constexpr int MAXBITS = %d;  // Scalar.maxBits\n" % maxbits
    s += "  static constexpr int SHIFT_K = %d;\n  static constexpr int SHIFT_M = %d;\n" % (k, m)
    s += "  static constexpr int QBITS = %d;\n" % ceil_log2(q)
    for nm, v in (("M0", m0), ("M1", m1)):
        s += "  static constexpr int %s_NEG = %d;\n" % (nm, 1 if v < 0 else 0)
        s += arr(nm, limbs(abs(v), 5))
    for nm, v in (("V00", v00), ("V01", v01), ("V10", v10), ("V11", v11)):
        s += "  static constexpr int %s_NEG = %d;\n" % (nm, 1 if v < 0 else 0)
        s += arr(nm, limbs(abs(v), 5))
    s += arr("Q", limbs(q, 8))
    s += "};\n\n"
    return s


def main():
    out = ("// GENERATED by tools/gen_constants.py -- do not edit.\n"
           "// 32-bit little-endian limbs; Montgomery radix R = 2^RBITS, RBITS = 32 N.\n"
           "#pragma once\n#include \"fp.cuh\"\n\nnamespace msm {\n\n")
    out += field_struct("Bls377Fq", BLS377_P, 12, {"BETA": BLS377_BETA, "GX": BLS377_GX, "GY": BLS377_GY})
    out += field_struct("PallasFp", PALLAS_P, 8, {"BETA": PALLAS_BETA, "GX": PALLAS_GX, "GY": PALLAS_GY})
    out += field_struct("Bls377Fr", ED_P, 8, {"K2D": 2 * ED_D % ED_P, "GX": ED_GX, "GY": ED_GY})
    out += field_struct("Bls381Fq", BLS381_P, 12, {"BETA": BLS381_BETA, "GX": BLS381_GX, "GY": BLS381_GY})
    out += glv_struct("Bls377Glv", BLS377_R, BLS377_LAMBDA)
    out += glv_struct("PallasGlv", PALLAS_Q, PALLAS_LAMBDA)
    out += glv_struct("Bls381Glv", BLS381_R, BLS381_LAMBDA)
    out += "struct EdScalar {\n  static constexpr int QBITS = %d;\n" % ceil_log2(ED_Q)
    out += arr("Q", limbs(ED_Q, 8)) + "};\n\n"
    out += "}  // namespace msm\n"
    with open(OUT, "w") as f:
        f.write(out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
