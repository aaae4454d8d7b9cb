// Pallas engine instantiation (src/concrete/pasta.params.ts; b = 5 -> 3b = 15).
#include "engine.cuh"
MSM_DEFINE_WEIERSTRASS_CURVE(curve_ops_pallas, PallasFp, PallasGlv, 15)
