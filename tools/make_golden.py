#!/usr/bin/env python3
"""Writes tests/golden/msm_vectors.json: small MSM input/output vectors computed by the python oracle
(oracle/bigint_oracle.py, the restatement of the reference's src/bigint/msm.ts), plus the reference's own
known-answer points (scripts/zprize23/submission-test-bls377.ts:6-25, submission-test.ts:5-21).
The reference itself cannot run in this image (no node), so these vectors are oracle outputs frozen at
the commit that pinned the oracle to the reference's KATs; tests compare BOTH the oracle and the GPU path
against them."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bigint_oracle as O  # noqa: E402

out = {"generator": "tools/make_golden.py", "cases": []}
for name, params in (("bls12-377", O.BLS12_377), ("pallas", O.PALLAS)):
    aff = O.WeierstrassAffine(params)
    for n, seed in ((1, 1), (5, 2), (32, 3)):
        pts = O.random_points_weierstrass(aff, n, seed)
        sc = O.random_scalars(n, params.q, seed + 100)
        res = O.msm(aff, sc, pts)
        out["cases"].append({"curve": name, "n": n, "points": [[hex(x), hex(y)] for x, y in pts],
                             "scalars": [hex(s) for s in sc], "result": [hex(res[0]), hex(res[1])]})
te = O.TwistedEdwards(O.ED_ON_BLS12_377)
for n, seed in ((1, 1), (5, 2), (32, 3)):
    pts = O.random_points_te(te, n, seed)
    sc = O.random_scalars(n, te.q, seed + 100)
    res = te.to_affine(O.msm(te, sc, [te.from_affine(p) for p in pts]))
    out["cases"].append({"curve": "ed-on-bls12-377", "n": n, "points": [[hex(x), hex(y)] for x, y in pts],
                         "scalars": [hex(s) for s in sc], "result": [hex(res[0]), hex(res[1])]})
out["kat"] = {
    "bls12-377": {"point": [hex(111871295567327857271108656266735188604298176728428155068227918632083036401841336689521497731900230387779623820740),
                            hex(76860045326390600098227152997486448974650822224305058012700629806287380625419427989664237630603922765089083164740)],
                  "scalars": ["0x2", hex(O.BLS12_377.q - 1)]},
    "ed-on-bls12-377": {"point": [hex(2796670805570508460920584878396618987767121022598342527208237783066948667246),
                                  hex(8134280397689638111748378379571739274369602049665521098046934931245960532166)],
                        "scalars": ["0x2", hex(O.ED_ON_BLS12_377.q - 1)]},
}
path = os.path.join(ROOT, "tests", "golden", "msm_vectors.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path, len(out["cases"]), "cases")
