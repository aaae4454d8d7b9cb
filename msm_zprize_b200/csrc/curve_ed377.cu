// ed-on-bls12-377 (twisted Edwards, a = -1) engine instantiation
// (src/concrete/ed-on-bls12-377.params.ts).
#include "engine.cuh"

static int run_te(msm_b200_ctx* ctx, const void* d_s, size_t n, int layout, int form, int c, msm_b200_timing* tm,
                  uint32_t* digits_dump_dev) {
  (void)form;
  (void)digits_dump_dev;
  return run_bucket_basic<TeCurve<Bls377Fr>, EdScalar>(ctx, d_s, n, layout, c, tm);
}

const CurveOps* curve_ops_ed377() {
  static const CurveOps ops = {ingest_te<Bls377Fr, EdScalar>,
                               run_te,
                               zero_partial_t<TeCurve<Bls377Fr>>,
                               finalize_any<TeCurve<Bls377Fr>>,
                               random_points_t<TeCurve<Bls377Fr>, EdScalar>,
                               random_scalars_t<EdScalar>};
  return &ops;
}
