#!/usr/bin/env python3
"""Prints the integer-pipe micro-benchmarks (the measured IMAD roofline denominators)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from msm_zprize_b200.engine import microbench  # noqa: E402

# SASS of every variant: profiles/r02_sass_hist_mb_imad_*.txt (0: pure IMAD; 1: ptxas splits every mad.wide.u32 with a
# 64-bit addend into IMAD.WIDE.U32 + IADD3 pairs, so its rate is the rate of that pair, not of IMAD.WIDE alone;
# 2: IMAD.WIDE.U32(.X) carry chains, the form the Montgomery product uses -- the roofline denominator)
NAMES = {0: "imad_lo", 1: "imad_wide_plus_iadd3", 2: "imad_wide_carry", 5: "imad_hi", 8: "iadd", 3: "modmul_12limb", 4: "modmul_8limb",
         6: "modsqr_12limb", 7: "modsqr_8limb", 9: "dbl_chain_quad_12limb", 10: "dbl_chain_lone_12limb",
         11: "dbl_chain_quad_8limb"}
out = {}
for which, name in NAMES.items():
    iters = 256
    ops, ms = microbench(0, which, iters)
    out[name] = {"ops_per_s": ops, "ms": ms}
    print("%-20s %10.3f Gop/s  (%.3f ms)" % (name, ops / 1e9, ms), flush=True)
out["modmul_12limb_LP_per_s"] = out["modmul_12limb"]["ops_per_s"] * 300
out["modmul_8limb_LP_per_s"] = out["modmul_8limb"]["ops_per_s"] * 136
print(json.dumps(out))
