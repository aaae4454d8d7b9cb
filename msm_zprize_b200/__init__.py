"""msm_zprize_b200 -- B200-native multi-scalar multiplication behind the msm entry points of
mitschabaude/msm-zprize (scripts/run-msm-377.ts, run-msm-pallas.ts, run-msm-ed-377.ts).

The CUDA engine lives in csrc/ (libmsm_b200.so, C ABI in include/msm_b200.h); this package is
the host-side mirror of the reference's `Parallel` interface (src/parallel.ts) over that ABI.
"""
from ._lib import (CURVE_BLS12_377_G1, CURVE_BLS12_381_G1, CURVE_ED_ON_BLS12_377, CURVE_PALLAS, FORM_AFFINE_GLV,
                   FORM_PROJECTIVE, FORM_TE_EXTENDED, LAYOUT_LE_BYTES, LAYOUT_LIMB29_MONT, MsmError)
from .engine import MsmEngine, MsmPipeline, MsmResult, MultiMsmEngine, PinnedBuffer, PipelinedMsm
from .submission import Submission

__all__ = ["MsmEngine", "MultiMsmEngine", "MsmPipeline", "PipelinedMsm", "MsmResult", "MsmError", "PinnedBuffer", "Submission", "CURVE_BLS12_377_G1", "CURVE_PALLAS",
           "CURVE_ED_ON_BLS12_377", "CURVE_BLS12_381_G1", "FORM_AFFINE_GLV", "FORM_PROJECTIVE", "FORM_TE_EXTENDED",
           "LAYOUT_LE_BYTES", "LAYOUT_LIMB29_MONT"]
