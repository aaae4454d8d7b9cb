// BLS12-381 G1 engine instantiation (src/concrete/bls12-381.params.ts; b = 4 -> 3b = 12).
// p != 1 mod 2^32 here, so the Montgomery factor M0 is a genuine multiplier.
#include "engine.cuh"
MSM_DEFINE_WEIERSTRASS_CURVE(curve_ops_bls381, Bls381Fq, Bls381Glv, 12)
