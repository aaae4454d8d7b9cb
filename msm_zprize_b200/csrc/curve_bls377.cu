// BLS12-377 G1 engine instantiation (src/concrete/bls12-377.params.ts; b = 1 -> 3b = 3).
#include "engine.cuh"
MSM_DEFINE_WEIERSTRASS_CURVE(curve_ops_bls377, Bls377Fq, Bls377Glv, 3)
