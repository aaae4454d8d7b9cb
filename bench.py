#!/usr/bin/env python3
"""bench.py -- headline benchmark of the MSM hot path (contract: see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on host cores

Workload (BASELINE.json configs[1]): BLS12-377 G1 MSM, n = 2^18 points per GPU, batched-affine GLV
path, synthetic seeded inputs (points built like randomPointsFast, scalars uniform in [0, q)).
A step = one MSM over the rank's point range with fresh scalars; bases stay resident
(scripts/msm-weierstrass.ts:12-51).  N > 1: every rank owns a contiguous range of an N * 2^18
point problem (weak scaling), partials are all-gathered over NCCL and summed on rank 0.

`value`  = Mpoints/s with inputs resident in HBM (CUDA events on the engine's stream).
`e2e`    = the same through the one-shot C-ABI call with pinned HOST buffers for scalars AND
           points (host->device copies, ingest and the result read-back inside the timed region).
`roofline` = the dominant kernel (batched-affine backward pass) against the IMAD.WIDE issue rate
           measured live by the library's micro-benchmark on the same GPU.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bls12_377_g1_msm_throughput"  # BASELINE.json's metric; other curves: <curve>_msm_throughput


def metric_name(curve):
    return METRIC if curve == "bls12-377" else curve.replace("-", "_") + "_msm_throughput"


# stdout carries exactly ONE line, the JSON result: libraries that print banners to fd 1 (NCCL's version line)
# are sent to stderr for the lifetime of the process
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())
UNIT = "Mpoints/s"
CURVES = {"bls12-377": (48, 12, 126), "pallas": (32, 8, 127), "ed-on-bls12-377": (32, 8, 251)}


def workload_name(curve, log2n):
    kind = "extended twisted-Edwards bucket method" if curve == "ed-on-bls12-377" else "batched-affine GLV"
    return f"{curve} MSM, n=2^{log2n} points per GPU, {kind}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--curve", default="bls12-377", choices=list(CURVES))
    ap.add_argument("--log2n", type=int, default=18, help="points per GPU = 2^log2n")
    ap.add_argument("--window", type=int, default=0)
    ap.add_argument("--cpu-log2n", type=int, default=15, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ------------------------------------------------------------------------------------------------
# algorithmic work: limb products of one MSM
# ------------------------------------------------------------------------------------------------
def w_alg(curve, n, c, te=False):
    nbytes, n32, b = CURVES[curve]
    lp = 2 * n32 * n32 + n32
    K = -(-(b + 1) // c)
    L = 1 << (c - 1)
    S = n if te else 2 * n
    A = S * K * (1 - 2.0 ** -c) - K * L * (1 - math.exp(-S / L))
    if te:
        return (8 * A + 18 * K * L) * lp
    return (6 * A + 25 * K * L + n) * lp


# ------------------------------------------------------------------------------------------------
# CPU arm: the C++ port of the reference's algorithm on the host cores (oracle/msm_port.cpp)
# ------------------------------------------------------------------------------------------------
def cpu_baseline(curve, points_le, scalars_le, n, threads, reps=1):
    from oracle.port import Port  # checker / baseline only -- never on the product path
    port = Port(curve)
    nb = CURVES[curve][0]
    prepared = port.prepare_points(points_le[: n * 2 * nb], n, threads)
    best = None
    res = None
    for _ in range(reps):
        x, y, z, sec = port.msm(scalars_le[: n * 32], prepared, n, threads)
        best = sec if best is None else min(best, sec)
        res = (x, y, z)
    return best, res, port.default_window(n)


def run_reference(args):
    """--impl reference: the reference's own algorithm (CPU port; the wasm original cannot run here)
    on all host cores, same metric / unit / config, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import msm_zprize_b200 as mz
    threads = os.cpu_count() or 1
    n = 1 << min(args.log2n, args.cpu_log2n + 1)
    nb = CURVES[args.curve][0]
    with mz.MsmEngine(args.curve) as eng:  # inputs come from the same seeded device generators
        d_pts = eng.dev_alloc(n * 2 * nb)
        d_sc = eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, n, 0xB200 + args.log2n)
        pts = eng.d2h(d_pts, n * 2 * nb).tobytes()
        scal = []
        for s in range(args.steps + args.warmup):
            eng.random_scalars_device(d_sc, n, 0x5CA1A + s)
            scal.append(eng.d2h(d_sc, n * 32).tobytes())
    from oracle.port import Port
    port = Port(args.curve)
    prepared = port.prepare_points(pts, n, threads)
    times = []
    for s in range(args.steps + args.warmup):
        _, _, _, sec = port.msm(scal[s], prepared, n, threads)
        if s >= args.warmup:
            times.append(sec)
    ms = 1e3 * sum(times) / len(times)
    value = n / (ms * 1e-3) / 1e6
    sample = f"2^{int(math.log2(n))} points per step (bounded sample of the 2^{args.log2n} workload), c={port.default_window(n)}"
    emit(({
        "impl": "reference", "metric": metric_name(args.curve), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.curve, args.log2n),
                   "reference_kind": "C++ port of the reference algorithm (29-bit-limb Montgomery, GLV, counting sort, "
                                     "batched affine), not the wasm original (no node in this image)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import msm_zprize_b200 as mz
    from msm_zprize_b200.engine import PinnedBuffer, microbench

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(dev)  # the engine launches everything on this stream
    torch.cuda.set_stream(stream)

    curve = args.curve
    nb, n32, _ = CURVES[curve]
    te = curve == "ed-on-bls12-377"
    n = 1 << args.log2n
    steps, warm = args.steps, max(args.warmup, 3)
    eng = mz.MsmEngine(curve, device=local, stream=stream.cuda_stream)

    # ---- synthetic inputs, generated on the device (seed = 0xB200 + log2n, per-rank ranges)
    pts = torch.empty(n * 2 * nb, dtype=torch.uint8, device=dev)
    eng.random_points_device(pts.data_ptr(), n, 0xB200 + args.log2n + 1000 * rank)
    eng.set_bases_device(pts.data_ptr(), n)
    total_steps = steps + warm
    scal = torch.empty((total_steps, n * 32), dtype=torch.uint8, device=dev)
    for s in range(total_steps):
        eng.random_scalars_device(scal[s].data_ptr(), n, 0x5CA1A + 7919 * s + 1000 * rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    partial = torch.zeros(eng.partial_bytes(), dtype=torch.uint8, device=dev)
    gathered = torch.zeros(world * eng.partial_bytes(), dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    class Step:  # timing of the rank's own MSM + the combined point (rank 0)
        def __init__(self, timing, res):
            self.timing, self.res = timing, res
            self.x, self.y = (res.x, res.y) if res is not None else (None, None)

    def step_resident(s):
        """one MSM over this rank's range, inputs resident in HBM"""
        if world == 1:
            r = eng.run(scal[s].data_ptr(), n, on_device=True, window_bits=args.window)
            return Step(r.timing, r)
        # no host synchronisation between the MSM, the collective and the combine: all three are ordered on
        # the engine's stream; the phase timings of the MSM are fetched afterwards
        eng.run_partial(scal[s].data_ptr(), n, partial.data_ptr(), on_device=True, window_bits=args.window, timing=False)
        dist.all_gather_into_tensor(gathered, partial)  # the only collective: world x 144 bytes
        res = eng.combine(gathered.data_ptr(), world) if rank == 0 else None
        return Step(eng.last_timing(), res)

    # ---- device-resident timing: K steps, per-step CUDA events on the engine's stream, L2 flushed
    #      between steps (outside the event pairs)
    for s in range(warm):
        step_resident(s)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    hot_ms = hot_launches = launches = n_adds = 0
    last = None
    clocks = ClockSampler(local)
    clocks.__enter__()
    if True:
        t_wall0 = time.perf_counter()
        for i in range(steps):
            flush.fill_(i & 0xFF)
            ev[i][0].record(stream)
            last = step_resident(warm + i)
            ev[i][1].record(stream)
            tm = last.timing
            hot_ms += tm["hot_kernel_ms"]
            hot_launches += tm["hot_kernel_launches"]
            launches += tm["kernel_launches"]
            n_adds += tm["n_adds"]
        barrier()
        wall_total = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_local = sum(ms_steps) / steps
    t = torch.tensor([ms_local], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item())
    value = world * n / (ms_per_step * 1e-3) / 1e6
    c_used, K_used = last.timing["window_bits"], last.timing["n_windows"]

    # ---- end to end: one-shot C-ABI call with pinned host buffers (points + scalars H2D, ingest,
    #      MSM, result D2H inside the timed region)
    h_pts = PinnedBuffer(n * 2 * nb)
    h_pts.array[:] = pts.cpu().numpy()
    h_sc = [PinnedBuffer(n * 32) for _ in range(total_steps)]
    for s in range(total_steps):
        h_sc[s].array[:] = scal[s].cpu().numpy()

    def step_e2e(s):
        if world == 1:
            return eng.msm(h_sc[s].array, h_pts.array, n, window_bits=args.window)
        eng.set_bases_async(h_pts.array, n)  # overlaps the scalar upload, GLV and sort of the run below
        eng.run_partial(h_sc[s].array, n, partial.data_ptr(), window_bits=args.window, timing=False)
        dist.all_gather_into_tensor(gathered, partial)
        return eng.combine(gathered.data_ptr(), world) if rank == 0 else None

    for s in range(warm):
        step_e2e(s)
    barrier()
    e2e_times = []
    for i in range(steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        r_e2e = step_e2e(warm + i)
        torch.cuda.synchronize(dev)
        e2e_times.append((time.perf_counter() - t0) * 1e3)
    barrier()
    clocks.__exit__()
    t = torch.tensor([sum(e2e_times) / steps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * n / (e2e_ms * 1e-3) / 1e6
    if rank == 0 and world == 1:
        assert (r_e2e.x, r_e2e.y) == (last.x, last.y), "e2e and resident paths disagree"

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel against the live-measured IMAD.WIDE issue rate
    peak_lp, _ = microbench(local, 2, 256)  # IMAD.WIDE.U32(.X): one 32x32->64 limb product each
    lp_mod = 2 * n32 * n32 + n32
    hot_modmul = (7 if te else 5)  # modmuls of one addition done inside the dominant kernel
    hot_lp = n_adds * hot_modmul * lp_mod
    achieved = hot_lp / (hot_ms * 1e-3) if hot_ms > 0 else 0.0
    # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/), when
    # it was taken on this workload; a live run cannot measure it without a profiler
    traffic, traffic_src = None, None
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))
        if rec["workload"] == workload_name(args.curve, args.log2n) and args.window == 0:
            traffic = rec["dram_bytes_per_launch"]
            traffic_src = "profiles/r01_ncu_traffic.json (dram__bytes_read.sum + dram__bytes_write.sum, mean of the " \
                          "%d launches of one step)" % rec["launches"]
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
        "algorithmic_work": f"{hot_modmul} of the {8 if te else 6} modmuls per point addition x {lp_mod} limb products x "
                            f"{n_adds // steps} additions per step",
        "launches_per_step": hot_launches // steps, "avg_launch_ms": hot_ms / max(hot_launches, 1),
        "share_of_step": hot_ms / (sum(ms_steps)) if world == 1 else None,
        "whole_msm_frac": (w_alg(curve, n, c_used, te) / (ms_per_step * 1e-3)) / peak_lp if peak_lp else None,
    }

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ncpu = 1 << min(args.cpu_log2n, args.log2n)
        threads = os.cpu_count() or 1
        pts_le = pts[: ncpu * 2 * nb].cpu().numpy().tobytes()
        sc_le = scal[warm][: ncpu * 32].cpu().numpy().tobytes()
        sec, res_cpu, c_cpu = cpu_baseline(curve, pts_le, sc_le, ncpu, threads)
        # the sample doubles as a parity check: GPU on the same prefix must give the same point
        eng.set_bases_device(pts.data_ptr(), n)
        chk = eng.run(scal[warm].data_ptr(), ncpu, on_device=True)
        parity = (chk.x, chk.y, chk.is_zero) == res_cpu
        cpu = {"value": ncpu / sec / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first 2^{int(math.log2(ncpu))} points/scalars of the workload, one run, "
                         f"reference window c={c_cpu}; GPU result on the same sample identical: {parity}",
               "ms": sec * 1e3}

    out = {
        "metric": metric_name(curve), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(curve, args.log2n), "window_bits": c_used, "n_windows": K_used,
                   "points_total": world * n, "l2": "flushed between timed steps (256 MiB write)",
                   "parallelism": f"range-sharded x{world}" if world > 1 else "single GPU",
                   "inputs": "seeded device generators (randomPointsFast construction, uniform scalars), bases resident"},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": n * (2 * nb + 32), "d2h_bytes_per_step": 100,
                "what": "msm_b200_msm(): pinned host scalars + points -> device, ingest, MSM, result read back"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "latency_ms": {"resident": ms_per_step, "e2e": e2e_ms, "steps_ms": [round(x, 3) for x in ms_steps]},
        "phases_ms_last_step": {k: round(last.timing[k], 3) for k in
                                ("digits_ms", "sort_ms", "accumulate_ms", "hot_kernel_ms", "reduce_ms")},
        "wall_s_timed_region": wall_total,
    }
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
