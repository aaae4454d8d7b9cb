#!/usr/bin/env python3
"""Opcode histogram of one kernel from `cuobjdump -sass` (evidence that the field product is IMAD.WIDE carry
chains and nothing else hides in the hot loops).

    python tools/sass_hist.py msm_zprize_b200/csrc/curve_bls377.o 'k_bwdINS_8Bls377FqELb0ELb1' > profiles/r02_sass_hist_k_bwd.txt
"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, hist, total = None, collections.Counter(), 0
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        if name is None and pat in cur:
            name = cur
        continue
    if cur != name or name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        hist[op] += 1
        total += 1
if name is None:
    sys.exit("no function matches " + pat)
print("# %s" % obj)
print("# function %s" % name)
print("# %d SASS instructions" % total)
groups = collections.Counter()
for op, c in hist.items():
    groups[op.split(".")[0]] += c
print("# by mnemonic:", ", ".join("%s %d" % (k, v) for k, v in groups.most_common(12)))
for op, c in hist.most_common():
    print("%-28s %6d  %5.1f %%" % (op, c, 100.0 * c / total))
