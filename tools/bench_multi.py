#!/usr/bin/env python3
"""In-library multi-GPU MSM (msm_b200_multi_*, no torch / torchrun): ONE process, one call per MSM.
Times the resident-shard path (scalars already on the devices) and the host path (scalars in pinned host
memory, range-sharded and uploaded by the library) for a fixed total of 2^T points over 1..G devices, and
checks every result against the single-device one.  dev tool; prints one JSON line per device count."""
import argparse
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msm_zprize_b200 as mz  # noqa: E402
from msm_zprize_b200.engine import PinnedBuffer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--curve", default="bls12-377")
ap.add_argument("--total-log2n", type=int, default=22)
ap.add_argument("--devices", default="1,2")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--depth", type=int, default=0, help="> 0: also time msm_b200_pipeline_* with that many lanes (host scalars)")
args = ap.parse_args()

n = 1 << args.total_log2n
ref = None
for G in [int(x) for x in args.devices.split(",")]:
    with mz.MultiMsmEngine(args.curve, list(range(G))) as m:
        pb = m.shards[0].point_bytes(mz.LAYOUT_LE_BYTES)
        per = -(-n // G)
        pp, ss, cnt = [], [], []
        for g, e in enumerate(m.shards):
            lo = min(n, per * g)
            k = min(n, lo + per) - lo
            p, s = e.dev_alloc(max(k, 1) * pb), e.dev_alloc(max(k, 1) * 32)
            e.random_points_device(p, k, 0xB200 + args.total_log2n, first=lo)
            e.random_scalars_device(s, k, 0x5CA1A, first=lo)
            pp.append(p), ss.append(s), cnt.append(k)
        t0 = time.perf_counter()
        m.set_bases_sharded(pp, cnt)
        t_bases = time.perf_counter() - t0
        # host copy of the scalars (pinned), for the msm_b200_multi_run path
        h = PinnedBuffer(n * 32)
        for g, e in enumerate(m.shards):
            lo = min(n, per * g)
            h.array[lo * 32:(lo + cnt[g]) * 32] = e.d2h(ss[g], cnt[g] * 32)
        res = None
        times_dev, times_host = [], []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = m.run_sharded(ss)
            t1 = time.perf_counter()
            rh = m.run(h.array, n)
            t2 = time.perf_counter()
            assert (rh.x, rh.y) == (res.x, res.y)
            if i >= args.warmup:
                times_dev.append((t1 - t0) * 1e3)
                times_host.append((t2 - t1) * 1e3)
        if ref is None:
            ref = (res.x, res.y)
        md, mh = statistics.median(times_dev), statistics.median(times_host)
        print(json.dumps({"curve": args.curve, "total_log2n": args.total_log2n, "devices": G, "gather": m.gather_kind,
                          "resident_ms": round(md, 3), "resident_sd": round(statistics.stdev(times_dev), 3),
                          "host_scalars_ms": round(mh, 3), "mpoints_s_resident": round(n / md / 1e3, 1),
                          "mpoints_s_host": round(n / mh / 1e3, 1), "set_bases_s": round(t_bases, 3),
                          "window_bits": res.timing["window_bits"], "shared_buckets": res.timing["shared_buckets"],
                          "same_point_as_first": (res.x, res.y) == ref,
                          "per_device_ms": [round(t["digits_ms"] + t["sort_ms"] + t["accumulate_ms"] + t["reduce_ms"], 3)
                                            for t in m.last_timings()]}), flush=True)
        if args.depth > 0:
            # several sharded MSMs in flight: every lane is a multi context over the same devices with its own NCCL
            # communicators; host scalars, plain submit / wait calls
            pts_host = b"".join(bytes(e.d2h(pp[g], cnt[g] * pb)) for g, e in enumerate(m.shards))
            with mz.MsmPipeline(args.curve, list(range(G)), depth=args.depth) as pipe:
                pipe.set_bases(pts_host, n)
                for _ in range(args.warmup):
                    pipe.wait(pipe.submit(h.array, n))
                t0 = time.perf_counter()
                tickets = [pipe.submit(h.array, n) for _ in range(min(args.steps, 2 * args.depth))]
                done = 0
                total = args.steps
                submitted = len(tickets)
                last = None
                while tickets:
                    last = pipe.wait(tickets.pop(0))
                    done += 1
                    if submitted < total:
                        tickets.append(pipe.submit(h.array, n))
                        submitted += 1
                ms = (time.perf_counter() - t0) * 1e3 / total
            print(json.dumps({"curve": args.curve, "total_log2n": args.total_log2n, "devices": G, "pipeline_depth": args.depth,
                              "host_scalars_ms_per_msm": round(ms, 3), "mpoints_s": round(n / ms / 1e3, 1),
                              "same_point": (last.x, last.y) == (res.x, res.y)}), flush=True)
        h.free()
