#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MSM hot path (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

Workload (BASELINE.json configs[1]): BLS12-377 G1 MSM, n = 2^18 points per GPU, batched-affine GLV
path, synthetic seeded inputs (points built like randomPointsFast, scalars uniform in [0, q)).
A step = one MSM over the rank's point range with fresh scalars; bases stay resident, like the
points that stay in wasm memory across the reference's benchmark iterations
(scripts/msm-weierstrass.ts:12-51).  Statistics follow the reference's protocol
(scripts/evaluate-util.ts:3-20): warm-up runs discarded, then median +- sample sd of the timed runs.

`value`  = Mpoints/s with inputs resident in HBM (CUDA events on the engine's stream), median step.
`e2e`    = the same through the reference-facing call msm_b200_run() with the scalars in pinned HOST
           memory: scalar upload, MSM and the result read-back inside the timed region, bases resident.
           `e2e.oneshot` = msm_b200_msm() with points AND scalars uploaded every call (compute_msm shape).
`roofline` = the dominant kernel (batched-affine backward pass) against the IMAD.WIDE issue rate
           measured live by the library's micro-benchmark on the same GPU.
N > 1 (torchrun): every rank owns a contiguous range of an N * 2^18 point problem (weak scaling),
partials are all-gathered over NCCL and summed on rank 0.  A `strong` block adds fixed-total problems
(BASELINE.json configs 4 and 5) range-sharded over the N ranks, next to the same problem on one GPU.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bls12_377_g1_msm_throughput"  # BASELINE.json's metric; other curves: <curve>_msm_throughput
UNIT = "Mpoints/s"
# curve -> (coordinate bytes, 32-bit limbs, scalar bits the windows cover)
CURVES = {"bls12-377": (48, 12, 126), "pallas": (32, 8, 127), "ed-on-bls12-377": (32, 8, 251), "bls12-381": (48, 12, 127)}
POINT_SEED = 0xB200
SCALAR_SEED = 0x5CA1A


def metric_name(curve):
    return METRIC if curve == "bls12-377" else curve.replace("-", "_") + "_msm_throughput"


# stdout carries exactly ONE line, the JSON result: libraries that print banners to fd 1 (NCCL's version line)
# are sent to stderr for the lifetime of the process
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


def workload_name(curve, log2n):
    kind = "extended twisted-Edwards bucket method" if curve == "ed-on-bls12-377" else "batched-affine GLV"
    return f"{curve} MSM, n=2^{log2n} points per GPU, {kind}"


def point_seed(log2n, rank):
    return POINT_SEED + log2n + 1000 * rank


def scalar_seed(step, rank):
    return SCALAR_SEED + 7919 * step + 1000 * rank


def med_sd(xs):
    """median and sample standard deviation, the reference's `median +- sd` (scripts/evaluate-util.ts:3-20)"""
    return statistics.median(xs), (statistics.stdev(xs) if len(xs) > 1 else 0.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bls12-377", choices=list(CURVES))
    ap.add_argument("--log2n", type=int, default=18, help="points per GPU = 2^log2n")
    ap.add_argument("--total-log2n", type=int, default=0,
                    help="strong scaling: fixed total of 2^T points range-sharded over the ranks (overrides --log2n)")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the fixed-total block (configs 4 and 5)")
    ap.add_argument("--strong-steps", type=int, default=5)
    ap.add_argument("--pipeline-depth", type=int, default=4, help="contexts of the `pipelined` block (MSMs in flight)")
    ap.add_argument("--cpu-log2n-max", type=int, default=20, help="largest MSM the CPU legs run (reference's ceiling: 2^20)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop = threading.Event()
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def bind_to_gpu_numa_node(index):
    """Pin this process (and the pinned buffers it allocates afterwards) to the CPUs next to its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# algorithmic work: limb products of one MSM
# ------------------------------------------------------------------------------------------------
def w_alg(curve, n, c, te=False, shared_buckets=False):
    nbytes, n32, b = CURVES[curve]
    lp = 2 * n32 * n32 + n32
    K = -(-(b + 1) // c)
    L = 1 << (c - 1)
    S = n if te else 2 * n
    if shared_buckets:  # every window adds into ONE set of L buckets (precomputed 2^(kc) G tables)
        A = S * K * (1 - 2.0 ** -c) - L * (1 - math.exp(-S * K / L))
        return (8 * A + 18 * L) * lp if te else (6 * A + 25 * L + n) * lp
    A = S * K * (1 - 2.0 ** -c) - K * L * (1 - math.exp(-S / L))
    if te:
        return (8 * A + 18 * K * L) * lp
    return (6 * A + 25 * K * L + n) * lp


# ------------------------------------------------------------------------------------------------
# CPU arm: the C++ port of the reference's algorithm on the host cores (oracle/msm_port.cpp).
# This process never loads libmsm_b200.so: the inputs come from the port's own seeded generators,
# which are byte-identical to the CUDA ones (tests/test_gpu_generators.py).
# ------------------------------------------------------------------------------------------------
def cpu_time_msm(port, scalars, prepared, n, threads, c, reps):
    times, res = [], None
    for _ in range(reps):
        x, y, z, sec = port.msm(scalars, prepared, n, threads, window_bits=c)
        times.append(sec)
        res = (x, y, z)
    return times, res


def run_reference(args):
    """--impl reference: the reference's own algorithm (CPU port; the wasm original cannot run here: no node)
    on all host cores, same metric / unit / config and the SAME n as our arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.port import Port  # test / baseline infrastructure; the only library this arm loads
    threads = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    log2n = args.log2n if not args.total_log2n else args.total_log2n
    want = log2n + (int(math.log2(world)) if not args.total_log2n else 0)  # the whole job's points
    lg = min(want, args.cpu_log2n_max)
    n = 1 << lg
    port = Port(args.curve)
    pts = port.random_points(n, point_seed(args.log2n, 0), threads)
    prepared = port.prepare_points(pts, n, threads)
    total = args.steps + args.warmup
    c_ref = port.default_window(n)
    # the CPU is not handicapped by the reference's table: one probe run per neighbouring window size
    probe = {}
    sc0 = port.random_scalars(n, scalar_seed(0, 0), threads)
    for c in range(max(c_ref - 2, 2), c_ref + 4):
        probe[c] = cpu_time_msm(port, sc0, prepared, n, threads, c, 1)[0][0]
    c_best = min(probe, key=probe.get)
    runs = {}
    for c in sorted({c_ref, c_best}):
        times = []
        for s in range(total):
            sc = port.random_scalars(n, scalar_seed(s, 0), threads)
            t, _ = cpu_time_msm(port, sc, prepared, n, threads, c, 1)
            if s >= args.warmup:
                times.append(t[0] * 1e3)
        med, sd = med_sd(times)
        runs[c] = {"window_bits": c, "ms_median": med, "ms_sd": sd, "ms_mean": sum(times) / len(times),
                   "mpoints_s": n / (med * 1e-3) / 1e6}
    best = min(runs.values(), key=lambda r: r["ms_median"])
    value = best["mpoints_s"]
    sample = (f"2^{lg} points per step" + ("" if lg == want else f" (bounded sample of the 2^{want}-point job)") +
              f", {args.steps} timed runs after {args.warmup} discarded, median; windows: reference table c={c_ref} "
              f"({runs[c_ref]['ms_median']:.1f} ms), best of the probe c={best['window_bits']} ({best['ms_median']:.1f} ms)")
    emit({
        "impl": "reference", "metric": metric_name(args.curve), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["ms_median"], "ms_sd": best["ms_sd"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.curve, args.log2n), "points_per_step": n,
                   "reference_kind": "C++ port of the reference algorithm (29-bit-limb Montgomery, GLV, counting sort, "
                                     "batched affine), not the wasm original (no node in this image)",
                   "inputs": "oracle/msm_port.cpp seeded generators (byte-identical to the CUDA generators)"},
        "windows": {"reference_table": runs[c_ref], "best": best, "probe_ms": {str(k): v * 1e3 for k, v in probe.items()}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    """what every measurement needs: torch device / stream, distributed handles"""
    pass


def setup():
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    g = Ctx()
    g.torch, g.world, g.rank, g.local = torch, world, rank, local
    g.cpus = bind_to_gpu_numa_node(local)
    g.dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        g.dist = dist
    g.dev = torch.device("cuda", local)
    torch.cuda.set_device(g.dev)
    g.stream = torch.cuda.Stream(g.dev)  # the engines launch everything on this stream
    torch.cuda.set_stream(g.stream)
    g.flush = torch.empty(256 << 20, dtype=torch.uint8, device=g.dev)  # > 126 MB L2
    return g


def barrier(g):
    if g.dist is not None:
        g.dist.barrier()
    g.torch.cuda.synchronize(g.dev)


def max_over_ranks(g, x):
    t = g.torch.tensor([x], dtype=g.torch.float64, device=g.dev)
    if g.dist is not None:
        g.dist.all_reduce(t, op=g.dist.ReduceOp.MAX)
    return float(t.item())


class Sharded:
    """One rank's share of an MSM: engine, resident bases, device scalars; step() = MSM (+ gather + combine)."""

    def __init__(self, g, curve, n_local, log2n_tag, n_scalar_sets, window=0, seed_rank=None):
        import msm_zprize_b200 as mz
        torch = g.torch
        self.g, self.curve, self.n, self.window = g, curve, n_local, window
        nb = CURVES[curve][0]
        self.nb = nb
        self.eng = mz.MsmEngine(curve, device=g.local, stream=g.stream.cuda_stream)
        r = g.rank if seed_rank is None else seed_rank
        self.pts = torch.empty(n_local * 2 * nb, dtype=torch.uint8, device=g.dev)
        self.eng.random_points_device(self.pts.data_ptr(), n_local, point_seed(log2n_tag, r))
        self.eng.set_bases_device(self.pts.data_ptr(), n_local)
        self.scal = torch.empty((n_scalar_sets, n_local * 32), dtype=torch.uint8, device=g.dev)
        for s in range(n_scalar_sets):
            self.eng.random_scalars_device(self.scal[s].data_ptr(), n_local, scalar_seed(s, r))
        self.partial = torch.zeros(self.eng.partial_bytes(), dtype=torch.uint8, device=g.dev)
        self.gathered = torch.zeros(g.world * self.eng.partial_bytes(), dtype=torch.uint8, device=g.dev)

    def step(self, s, solo=False):
        """returns (timing dict of this rank's MSM, combined MsmResult on rank 0 or None)"""
        g, eng = self.g, self.eng
        sp = self.scal[s % self.scal.shape[0]].data_ptr()
        if g.world == 1 or solo:
            r = eng.run(sp, self.n, on_device=True, window_bits=self.window)
            return r.timing, r
        # no host synchronisation between the MSM, the collective and the combine: all three are ordered on
        # the engine's stream; the phase timings of the MSM are fetched afterwards
        eng.run_partial(sp, self.n, self.partial.data_ptr(), on_device=True, window_bits=self.window, timing=False)
        g.dist.all_gather_into_tensor(self.gathered, self.partial)  # the only collective: world x 144 bytes
        res = eng.combine(self.gathered.data_ptr(), g.world) if g.rank == 0 else None
        return eng.last_timing(), res

    def timed(self, steps, warm, first=0, solo=False):
        """`steps` timed MSMs after `warm` discarded ones; per-step CUDA events on the engine's stream, L2 flushed
        between steps (outside the event pairs).  Returns (list of ms, accumulated timing, last result)."""
        g, torch = self.g, self.g.torch
        for s in range(warm):
            self.step(first + s, solo)
        if not solo:
            barrier(g)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        acc = {"hot_kernel_ms": 0.0, "hot_kernel_launches": 0, "kernel_launches": 0, "n_adds": 0, "fwd_round0_ms": 0.0,
               "fwd_round0_pairs": 0}
        last_tm = last_res = None
        for i in range(steps):
            g.flush.fill_(i & 0xFF)
            ev[i][0].record(g.stream)
            last_tm, last_res = self.step(first + warm + i, solo)
            ev[i][1].record(g.stream)
            for k in acc:
                acc[k] += last_tm[k]
        if not solo:
            barrier(g)
        else:
            torch.cuda.synchronize(g.dev)
        return [a.elapsed_time(b) for a, b in ev], acc, last_tm, last_res

    def close(self):
        self.eng.close()
        del self.pts, self.scal, self.partial, self.gathered
        self.g.torch.cuda.empty_cache()


def pipelined_block(g, sh, curve, n, steps, warm, h_sc, depth=2):
    """`depth` contexts over ONE resident point set (msm_b200_share_bases), each driven by its own host thread: the
    same MSMs as the timed steps, submitted back to back so that the latency-bound phases of one overlap the rounds of
    another.  Whole-run wall clock between device-wide synchronisations; points compared with the sequential run."""
    import threading
    import msm_zprize_b200 as mz
    torch = g.torch
    engines = [sh.eng] + [mz.MsmEngine(curve, device=g.local) for _ in range(depth - 1)]
    for e in engines[1:]:
        e.share_bases(sh.eng)
    total = warm + steps
    want = {s: sh.eng.run(sh.scal[s].data_ptr(), n, on_device=True) for s in (warm, total - 1)}

    def run_all(host):
        res = [None] * total

        def one(i, s):
            res[s] = engines[i].run(h_sc[s].array, n) if host else engines[i].run(sh.scal[s].data_ptr(), n, on_device=True)

        def phase(lo, hi):  # steps lo .. hi - 1, round-robin over the contexts, one host thread per context
            ths = [threading.Thread(target=lambda i=i: [one(i, s) for s in range(lo + i, hi, depth)]) for i in range(depth)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()

        phase(0, warm)  # untimed
        torch.cuda.synchronize(g.dev)
        t0 = time.perf_counter()
        phase(warm, total)
        torch.cuda.synchronize(g.dev)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        return ms, all((res[s].x, res[s].y) == (want[s].x, want[s].y) for s in want)

    ms_res, ok1 = run_all(False)
    ms_host, ok2 = run_all(True)
    for e in engines[1:]:
        e.close()
    # the same through the plain-call pipeline of the C ABI (msm_b200_pipeline_submit / _wait: the lanes' host threads
    # live inside the library), pinned host scalars
    capi_ms, ok3 = None, True
    try:
        with mz.MsmPipeline(curve, [g.local], depth=depth) as pipe:
            pipe.set_bases(sh.pts.cpu().numpy(), n)
            for s in range(0, warm, depth):
                [pipe.wait(t) for t in [pipe.submit(h_sc[u].array, n) for u in range(s, min(warm, s + depth))]]
            torch.cuda.synchronize(g.dev)
            t0 = time.perf_counter()
            tickets = []
            res = {}
            for s in range(warm, total):
                tickets.append((s, pipe.submit(h_sc[s].array, n)))
                if len(tickets) >= 2 * depth:  # keep the ticket ring short: collect the oldest
                    s0, t = tickets.pop(0)
                    res[s0] = pipe.wait(t)
            for s0, t in tickets:
                res[s0] = pipe.wait(t)
            capi_ms = (time.perf_counter() - t0) * 1e3 / steps
            ok3 = all((res[s].x, res[s].y) == (want[s].x, want[s].y) for s in want)
    except Exception as e:
        capi_ms = repr(e)[:200]
    return {"contexts": depth, "c_api_e2e_ms_per_msm": capi_ms, "c_api_same_points": ok3, "ms_per_msm": ms_res, "value": n / (ms_res * 1e-3) / 1e6, "unit": UNIT,
            "e2e_ms_per_msm": ms_host, "e2e_value": n / (ms_host * 1e-3) / 1e6, "same_points_as_sequential": ok1 and ok2,
            "what": f"{depth} contexts sharing the resident bases, one host thread each, {steps} MSMs submitted back to back "
                    "(wall clock of the whole run / MSMs, no L2 flush in between); one MSM alone takes longer than in the "
                    "sequential steps, which stay the headline"}


def check_combined_point(g, sh, curve, s):
    """N > 1: rank 0's combined point must equal the sum of the ranks' own single-GPU results (added here with
    the python oracle, the checker)."""
    _, res = sh.step(s)
    own = sh.eng.run(sh.scal[s % sh.scal.shape[0]].data_ptr(), sh.n, on_device=True)
    objs = [None] * g.world
    g.dist.all_gather_object(objs, (own.x, own.y, own.is_zero))
    if g.rank != 0:
        return True
    from oracle import bigint_oracle as O
    if curve == "ed-on-bls12-377":
        te = O.TwistedEdwards(O.ED_ON_BLS12_377)
        acc = te.from_affine((0, 1))
        for x, y, _ in objs:
            acc = te.add(acc, te.from_affine((x, y)))
        want = te.to_affine(acc)
        return (res.x, res.y) == tuple(want)
    prm = {"bls12-377": O.BLS12_377, "pallas": O.PALLAS, "bls12-381": O.BLS12_381}[curve]
    aff = O.WeierstrassAffine(prm)
    acc = None  # the point at infinity
    for x, y, z in objs:
        if not z:
            acc = aff.add(acc, (x, y))
    if res.is_zero or acc is None:
        return bool(res.is_zero) and acc is None
    return (res.x, res.y) == acc


def strong_case(g, curve, T, steps, peak_lp, cpu=False):
    """Fixed total of 2^T points range-sharded over the ranks (BASELINE.json configs 4 / 5), and the same
    problem on ONE GPU (rank 0, the others wait) for the strong-scaling efficiency.  `cpu`: also time the CPU
    port on the same inputs (N = 1 only, sizes the reference itself can run) and compare the points."""
    lw = int(math.log2(g.world))
    te = curve == "ed-on-bls12-377"
    out = {"curve": curve, "total_log2n": T, "n_gpus": g.world, "points_per_gpu": 1 << (T - lw)}
    sh = Sharded(g, curve, 1 << (T - lw), T, 2)
    ms, acc, tm, res = sh.timed(steps, 2)
    med, sd = med_sd(ms)
    med = max_over_ranks(g, med)
    out.update({"ms": med, "ms_sd": sd, "mpoints_s": (1 << T) / (med * 1e-3) / 1e6, "window_bits": tm["window_bits"],
                "shared_buckets": tm["shared_buckets"],
                "whole_msm_frac_per_gpu": w_alg(curve, 1 << (T - lw), tm["window_bits"], te, bool(tm["shared_buckets"])) /
                (med * 1e-3) / peak_lp if peak_lp else None})
    if cpu and g.world == 1:
        from oracle.port import Port  # baseline / checker only
        n, nb, threads = 1 << T, CURVES[curve][0], os.cpu_count() or 1
        port = Port(curve)
        prepared = port.prepare_points(sh.pts.cpu().numpy().tobytes(), n, threads)
        s_last = (2 + steps - 1) % sh.scal.shape[0]  # the scalars of the last timed step
        times, res_cpu = cpu_time_msm(port, sh.scal[s_last].cpu().numpy().tobytes(), prepared, n, threads,
                                      port.default_window(n), 2)
        out["cpu_port_ms"] = min(times) * 1e3
        out["cpu_port_window_bits"] = port.default_window(n)
        out["cpu_port_cores"] = threads
        out["same_point_as_cpu_port"] = (res.x, res.y, res.is_zero) == res_cpu
    sh.close()
    if g.world > 1:
        single = None
        if g.rank == 0:
            s1 = Sharded(g, curve, 1 << T, T, 2, seed_rank=0)
            ms1, _, tm1, _ = s1.timed(max(3, steps // 2), 1, solo=True)
            single = med_sd(ms1)[0]
            s1.close()
        barrier(g)
        if g.rank == 0:
            out["single_gpu_ms"] = single
            out["speedup"] = single / med
            out["efficiency"] = single / med / g.world
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    g = setup()
    torch, dist, world, rank, local, dev, stream = g.torch, g.dist, g.world, g.rank, g.local, g.dev, g.stream
    from msm_zprize_b200.engine import PinnedBuffer, microbench

    curve = args.curve
    nb, n32, _ = CURVES[curve]
    te = curve == "ed-on-bls12-377"
    if args.total_log2n:
        args.log2n = args.total_log2n - int(math.log2(world))
    n = 1 << args.log2n
    steps, warm = args.steps, max(args.warmup, 3)
    total_steps = steps + warm

    # ---- synthetic inputs, generated on the device (seed = 0xB200 + log2n, per-rank ranges)
    sh = Sharded(g, curve, n, args.log2n, total_steps, args.window)
    eng = sh.eng

    # ---- device-resident timing
    clocks = ClockSampler(local)
    clocks.__enter__()
    t_wall0 = time.perf_counter()
    ms_steps, acc, last_tm, last = sh.timed(steps, warm)
    wall_total = time.perf_counter() - t_wall0
    med_local, sd_local = med_sd(ms_steps)
    ms_per_step = max_over_ranks(g, med_local)
    value = world * n / (ms_per_step * 1e-3) / 1e6
    c_used, K_used = last_tm["window_bits"], last_tm["n_windows"]
    hot_ms, hot_launches, launches, n_adds = (acc["hot_kernel_ms"], acc["hot_kernel_launches"], acc["kernel_launches"],
                                              acc["n_adds"])

    # ---- end to end through the reference-facing call: scalars in pinned host memory, bases resident
    h_sc = [PinnedBuffer(n * 32) for _ in range(total_steps)]
    for s in range(total_steps):
        h_sc[s].array[:] = sh.scal[s].cpu().numpy()

    def step_e2e(s):
        if world == 1:
            return eng.run(h_sc[s].array, n, window_bits=args.window)
        eng.run_partial(h_sc[s].array, n, sh.partial.data_ptr(), window_bits=args.window, timing=False)
        dist.all_gather_into_tensor(sh.gathered, sh.partial)
        return eng.combine(sh.gathered.data_ptr(), world) if rank == 0 else None

    def time_host_calls(fn):
        for s in range(warm):
            fn(s)
        barrier(g)
        ts, r = [], None
        for i in range(steps):
            g.flush.fill_(i & 0xFF)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            r = fn(warm + i)
            torch.cuda.synchronize(dev)
            ts.append((time.perf_counter() - t0) * 1e3)
        barrier(g)
        return ts, r

    e2e_times, r_e2e = time_host_calls(step_e2e)
    e2e_med_local, e2e_sd = med_sd(e2e_times)
    e2e_ms = max_over_ranks(g, e2e_med_local)
    e2e_value = world * n / (e2e_ms * 1e-3) / 1e6

    # one-shot shape (compute_msm): points AND scalars uploaded and ingested inside every call
    h_pts = PinnedBuffer(n * 2 * nb)
    h_pts.array[:] = sh.pts.cpu().numpy()

    def step_oneshot(s):
        if world == 1:
            return eng.msm(h_sc[s].array, h_pts.array, n, window_bits=args.window)
        eng.set_bases_async(h_pts.array, n)  # overlaps the scalar upload, GLV and sort of the run below
        eng.run_partial(h_sc[s].array, n, sh.partial.data_ptr(), window_bits=args.window, timing=False)
        dist.all_gather_into_tensor(sh.gathered, sh.partial)
        return eng.combine(sh.gathered.data_ptr(), world) if rank == 0 else None

    one_times, r_one = time_host_calls(step_oneshot)
    one_ms = max_over_ranks(g, med_sd(one_times)[0])
    eng.set_bases_device(sh.pts.data_ptr(), n)  # the one-shot calls left bases without window tables: resident shape again
    clocks.__exit__()
    ok = True
    notes = []
    if rank == 0:
        same = (r_e2e.x, r_e2e.y) == (last.x, last.y) and (r_one.x, r_one.y) == (last.x, last.y)
        if not same:
            ok = False
            notes.append("resident, e2e and one-shot paths disagree")
    if world > 1:
        if not check_combined_point(g, sh, curve, warm):
            ok = False
            notes.append("combined multi-GPU point differs from the sum of the per-rank results")

    # ---- several MSMs in flight over the same resident bases (throughput of a stream of MSMs; N = 1)
    pipelined = None
    if world == 1:
        try:
            pipelined = pipelined_block(g, sh, curve, n, steps, warm, h_sc, max(2, args.pipeline_depth))
            if not (pipelined["same_points_as_sequential"] and pipelined["c_api_same_points"]):
                ok = False
                notes.append("pipelined MSMs differ from the sequential ones")
        except Exception as e:  # an extra: must not take the headline line with it
            pipelined = {"error": repr(e)[:300]}

    # ---- roofline of the dominant kernel against the live-measured IMAD.WIDE issue rate
    peak_lp, _ = microbench(local, 2, 256)  # IMAD.WIDE.U32(.X): one 32x32->64 limb product each
    peak_imad, _ = microbench(local, 0, 256)  # plain IMAD (mad.lo), for the record
    shared = bool(last_tm.get("shared_buckets", 0))

    # ---- fixed-total problems sharded over the ranks (configs 4 and 5)
    strong = None
    if not args.no_strong and curve == "bls12-377" and not args.total_log2n and args.log2n == 18:
        strong = []
        # config 5: BLS12-377 sweep 2^20 .. 2^26; config 4: ed-on-bls12-377 2^22
        for cv, T in (("bls12-377", 20), ("bls12-377", 22), ("bls12-377", 24), ("bls12-377", 26), ("ed-on-bls12-377", 22)):
            try:
                strong.append(strong_case(g, cv, T, args.strong_steps, peak_lp, cpu=(T == 20)))
            except Exception as e:  # a failed extra must not take the headline line with it
                strong.append({"curve": cv, "total_log2n": T, "error": repr(e)[:300]})
                barrier(g)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    lp_mod = 2 * n32 * n32 + n32
    hot_modmul = (7 if te else 5)  # modmuls of one addition done inside the dominant kernel
    hot_lp = n_adds * hot_modmul * lp_mod
    achieved = hot_lp / (hot_ms * 1e-3) if hot_ms > 0 else 0.0
    # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/), when
    # it was taken on this workload; a live run cannot measure it without a profiler
    traffic, traffic_src = None, None
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", name)))
            if rec["workload"] == workload_name(curve, args.log2n) and args.window == 0:
                traffic = rec["dram_bytes_per_launch"]
                traffic_src = "profiles/%s (dram__bytes_read.sum + dram__bytes_write.sum, mean of the %d launches of " \
                              "one step)" % (name, rec["launches"])
                break
        except (OSError, KeyError, ValueError):
            pass
    roofline = {
        "bound": "imad", "kernel": "k_bucket_acc" if te else "k_bwd (batched-affine backward pass)",
        "achieved": achieved / 1e12, "peak": peak_lp / 1e12, "unit": "T limb-products/s",
        "frac": achieved / peak_lp if peak_lp else None, "traffic": traffic,
        "traffic_source": traffic_src,
        # bytes one addition has to move in this kernel: two points + the prefix product in, one point out, keys
        # (twisted Edwards: one cached base point + its entry; the accumulator stays in registers)
        "algorithmic_bytes": (n_adds / max(hot_launches, 1)) * ((3 * n32 * 4 + 4) if te else (7 * n32 * 4 + 16)),
        "peak_source": "live micro-benchmark: IMAD.WIDE.U32(.X) carry-chain issue rate on this GPU (msm_b200_microbench 2)",
        "pure_imad_peak": peak_imad / 1e12,
        "algorithmic_work": f"{hot_modmul} of the {8 if te else 6} modmuls per point addition x {lp_mod} limb products x "
                            f"{n_adds // steps} additions per step",
        "launches_per_step": hot_launches // steps, "avg_launch_ms": hot_ms / max(hot_launches, 1),
        "share_of_step": hot_ms / (sum(ms_steps)) if world == 1 else None,
        "whole_msm_frac": (w_alg(curve, n, c_used, te, shared) / (ms_per_step * 1e-3)) / peak_lp if peak_lp else None,
    }

    # secondary roofline: the round-0 forward pass, the HBM-bound kernel of the path (random gathers of the x
    # coordinates out of the resident record sets), against the measured copy bandwidth of MEASURED_PEAKS.json
    if not te and acc["fwd_round0_ms"] > 0:
        try:
            hbm_peak, hbm_src = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json (measured copy)"
        except (OSError, KeyError, ValueError):
            hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
        fb = 4 * n32
        bytes_pair = 2 * fb + fb + 8 + 4 + 8  # two x coordinates in, one prefix product out, entries, pair key, offsets
        fwd_gbs = acc["fwd_round0_pairs"] * bytes_pair / (acc["fwd_round0_ms"] * 1e-3) / 1e9
        ncu_bytes = None
        try:
            rec = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")))
            if rec["workload"] == workload_name(curve, args.log2n) and args.window == 0:
                ncu_bytes = rec["k_fwd_per_launch"][0]["dram_bytes"]
        except (OSError, KeyError, ValueError, IndexError):
            pass
        roofline["secondary"] = {
            "bound": "hbm", "kernel": "k_fwd round 0 (forward prefix products over gathered base points)",
            "achieved": fwd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": fwd_gbs / hbm_peak,
            "avg_launch_ms": acc["fwd_round0_ms"] / steps, "algorithmic_bytes": acc["fwd_round0_pairs"] / steps * bytes_pair,
            "traffic": ncu_bytes, "peak_source": hbm_src,
            "note": "48-byte gathers out of 96-byte records move whole 64-byte lines: the ncu capture shows about 2.2x the "
                    "algorithmic bytes, i.e. the DRAM side runs at about twice the fraction quoted here"}

    # ---- CPU baseline on the SAME workload (rank 0, N = 1 only): the port on all host cores, and the parity
    #      check of the GPU result at full size
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle.port import Port  # checker / baseline only -- never on the product path
        lg = min(args.log2n, args.cpu_log2n_max)
        ncpu = 1 << lg
        threads = os.cpu_count() or 1
        port = Port(curve)
        prepared = port.prepare_points(sh.pts[: ncpu * 2 * nb].cpu().numpy().tobytes(), ncpu, threads)
        sc_le = sh.scal[warm][: ncpu * 32].cpu().numpy().tobytes()
        c_cpu = port.default_window(ncpu)
        reps = 5 if lg <= 18 else 2
        times, res_cpu = cpu_time_msm(port, sc_le, prepared, ncpu, threads, c_cpu, reps)
        chk = eng.run(sh.scal[warm].data_ptr(), ncpu, on_device=True)
        parity = (chk.x, chk.y, chk.is_zero) == res_cpu
        if not parity:
            ok = False
            notes.append("GPU result differs from the CPU port on the same inputs")
        med = statistics.median(times)
        cpu = {"value": ncpu / med / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"the same 2^{lg} points/scalars as one GPU step, {reps} runs, median, reference window "
                         f"c={c_cpu}; GPU result on the same inputs identical: {parity}",
               "ms": med * 1e3}

    out = {
        "metric": metric_name(curve), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_per_step, "ms_sd": sd_local, "ms_mean": sum(ms_steps) / steps, "higher_is_better": True,
        "scaling": "strong" if args.total_log2n else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "ok": ok, "notes": notes,
        "config": {"workload": workload_name(curve, args.log2n), "window_bits": c_used, "n_windows": K_used,
                   "shared_buckets": shared,
                   "points_total": world * n, "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": f"range-sharded x{world}" if world > 1 else "single GPU",
                   "statistic": "median of the timed steps (sample sd in ms_sd), max over ranks",
                   "cpu_affinity": f"{len(g.cpus)} CPUs next to the GPU" if g.cpus else "unchanged",
                   "inputs": "seeded device generators (randomPointsFast construction, uniform scalars), bases resident"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "ms_sd": e2e_sd,
                "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 100,
                "what": "msm_b200_run(): pinned host scalars -> device, MSM over the resident bases, result read back "
                        "(the reference's benchmark shape: points stay in memory, fresh scalars per run)",
                "oneshot": {"ms_per_step": one_ms, "value": world * n / (one_ms * 1e-3) / 1e6,
                            "h2d_bytes_per_step": n * (2 * nb + 32),
                            "what": "msm_b200_msm(): points AND scalars uploaded and ingested inside every call"}},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "latency_ms": {"resident": ms_per_step, "e2e": e2e_ms, "oneshot": one_ms,
                       "steps_ms": [round(x, 3) for x in ms_steps]},
        "phases_ms_last_step": {k: round(last_tm[k], 3) for k in
                                ("digits_ms", "sort_ms", "accumulate_ms", "hot_kernel_ms", "reduce_ms")},
        "wall_s_timed_region": wall_total,
    }
    if pipelined is not None:
        out["pipelined"] = pipelined
    if strong is not None:
        out["strong"] = strong
        if any(c.get("same_point_as_cpu_port") is False for c in strong):
            out["ok"] = ok = False
            out["notes"].append("a strong-block GPU result differs from the CPU port")
    emit(out)
    if dist is not None:
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


if __name__ == "__main__":
    main()
