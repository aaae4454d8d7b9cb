#!/usr/bin/env python3
"""Randomised differential test of the multi-GPU and pipeline layers (msm_b200_multi_*, msm_b200_pipeline_*) against
the CPU port, for a given number of seconds on whatever GPUs are visible: random curve, size (incl. fewer points than
devices), device subset and order, gather kind, resident / one-shot / prefix runs, several MSMs in flight.
dev tool; the checker is oracle/ (test infrastructure)."""
import os
import random
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402
from oracle.port import Port  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = random.Random(seed)
threads = os.cpu_count() or 1
ngpu = sum(1 for l in subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.splitlines() if l.startswith("GPU "))
CURVES = ["bls12-377", "pallas", "ed-on-bls12-377"]
ports = {c: Port(c) for c in CURVES}
t_end = time.time() + seconds
cases = bad = 0
while time.time() < t_end:
    curve = rng.choice(CURVES)
    port = ports[curve]
    nb = port.nbytes
    k = rng.randint(1, ngpu)
    devices = rng.sample(range(ngpu), k)
    if rng.random() < 0.3:
        os.environ["MSM_B200_GATHER"] = "peer"
    else:
        os.environ.pop("MSM_B200_GATHER", None)
    lg = rng.choice([0, 1, 2, 5, 9, 12, 14, 15, 16])
    n = max(1, rng.randrange(1 << lg, (2 << lg)))
    pts = port.random_points(n, rng.getrandbits(40), threads)
    scs = [port.random_scalars(n, rng.getrandbits(40), threads) for _ in range(3)]
    prep = port.prepare_points(pts, n, threads)
    want = [port.msm(s, prep, n, threads)[:3] for s in scs]
    mode = rng.choice(["multi", "multi", "pipeline"])
    got = []
    if mode == "multi":
        with mz.MultiMsmEngine(curve, devices) as m:
            m.set_bases(pts, n)
            got.append(m.run(scs[0], n))
            got.append(m.msm(scs[1], pts, n))
            m.set_bases(pts, n)
            got.append(m.run(scs[2], n))
            pre = rng.randrange(0, n + 1)  # a prefix of the resident set, possibly empty
            r = m.run(scs[0], pre)
            w = port.msm(scs[0][:32 * pre], port.prepare_points(pts[:2 * nb * pre], pre, threads), pre, threads)[:3] if pre else None
            if pre and (r.x, r.y, r.is_zero) != w:
                bad += 1
                print("MISMATCH prefix", curve, n, pre, devices, flush=True)
    else:
        with mz.MsmPipeline(curve, devices, depth=rng.randint(1, 3)) as p:
            p.set_bases(pts, n)
            tickets = [p.submit(s, n) for s in scs]
            got = [p.wait(t) for t in tickets]
    cases += 1
    for i, r in enumerate(got):
        if (r.x, r.y, r.is_zero) != want[i]:
            bad += 1
            print("MISMATCH", mode, curve, "n", n, "devices", devices, "gather", os.environ.get("MSM_B200_GATHER"), "call", i, flush=True)
print("fuzz_multi: %d cases on %d GPUs, %d mismatches, seed %d, %.0f s" % (cases, ngpu, bad, seed, seconds))
sys.exit(1 if bad else 0)
