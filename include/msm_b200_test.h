/* msm_b200_test.h -- test and measurement hooks of libmsm_b200.so.  Not part of the drop-in boundary
 * (include/msm_b200.h); used by tests/, tools/ and bench.py's roofline leg only. */
#ifndef MSM_B200_TEST_H
#define MSM_B200_TEST_H

#include "msm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/*  * Field ops on the device, element-wise over `n` elements (32-bit limbs, internal Montgomery
 * form) -- the device side of the reference's per-op tests (src/field.test.ts:15-155).
 * field: 0 BLS12-377 Fq, 1 Pallas Fp, 2 BLS12-377 Fr, 3 BLS12-381 Fq.  op: 0 mul, 1 add, 2 sub, 3 inverse,
 * 4 square, 5 inverse by the quad-cooperative routine (csrc/inv_quad.cuh). */
int msm_b200_test_field_op(int device, int field, int op, const uint32_t* a_host,
                           const uint32_t* b_host, uint32_t* out_host, size_t n);
/* GLV decomposition + signed digits on the device for `n` scalars (LE_BYTES): writes
 * 2n * K digits as u32 (bucket l | sign << 31), half-scalar major.  (src/glv/glv-test.ts,
 * src/msm-batched-affine.ts:172-200) */
int msm_b200_test_digits(msm_b200_ctx* ctx, const void* scalars_host, size_t n, int window_bits,
                         uint32_t* digits_host, int* n_windows);
/* Integer-pipe micro-benchmarks (the measured roofline denominators): which = 0 IMAD (mad.lo),
 * 1 IMAD.WIDE (mad.wide.u32), 2 IMAD.WIDE with carry in/out (mad.lo.cc/madc.hi.cc chains),
 * 5 IMAD.HI, 8 IADD3, 3 Montgomery product 12 limbs, 4 Montgomery product 8 limbs,
 * 6 / 7 Montgomery squaring 12 / 8 limbs; 9 / 10 / 11 latency of a chain of projective doublings in ONE warp
 * (quad-cooperative 12 limbs / one lane 12 limbs / quad-cooperative 8 limbs: what bounds the Horner tail).
 * Returns operations per second (limb products for 0-2 and 5, adds for 8, modmuls for 3-4 and 6-7,
 * doublings for 9-11). */
int msm_b200_microbench(int device, int which, int iters, double* ops_per_sec, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_TEST_H */
