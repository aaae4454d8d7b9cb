"""Multi-process path on the CPU (gloo, world_size 2): range sharding + gather of partials.
Each rank computes the MSM of its range with the C++ port (the checker standing in for the GPU
engine, which needs a device), the partials are all-gathered with the same helper the GPU path
uses, and rank 0's sum must equal the MSM over the whole range."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bigint_oracle as O


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from msm_zprize_b200.dist import gather_partials, shard_range
        from oracle.port import Port
        from tests import inputs as I
        aff = O.WeierstrassAffine(O.BLS12_377)
        pts = O.random_points_weierstrass(aff, n, seed=5)
        sc = O.random_scalars(n, aff.q, seed=6)
        lo, hi = shard_range(n, rank, world)
        p = Port("bls12-377")
        prep = p.prepare_points(I.points_le(pts[lo:hi], 48), hi - lo)
        x, y, z, _ = p.msm(I.scalars_le(sc[lo:hi]), prep, hi - lo, 1, 4)
        blob = x.to_bytes(48, "little") + y.to_bytes(48, "little") + bytes([1 if z else 0]) + bytes(47)
        allp = gather_partials(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        if rank == 0:
            acc = None
            raw = allp.numpy().tobytes()
            for r in range(world):
                b = raw[r * 144:(r + 1) * 144]
                P = None if b[96] else (int.from_bytes(b[:48], "little"), int.from_bytes(b[48:96], "little"))
                acc = aff.add(acc, P)
            q.put(acc == O.msm(aff, sc, pts))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from msm_zprize_b200.dist import shard_range
    for n in (0, 1, 7, 8, 1000, 1 << 18):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_library_shard_range_equals_the_harness():
    """msm_b200_multi_* (one process, all GPUs) and the torchrun harness must cut a point set the same way."""
    import ctypes as C
    from msm_zprize_b200 import _lib
    from msm_zprize_b200.dist import shard_range
    L = _lib.lib()
    for n in (0, 1, 7, 8, 1000, (1 << 18) + 3, 1 << 26):
        for world in (1, 2, 3, 4, 8):
            for r in range(world):
                lo, cnt = C.c_size_t(), C.c_size_t()
                L.msm_b200_multi_shard_range(n, r, world, C.byref(lo), C.byref(cnt))
                a, b = shard_range(n, r, world)
                assert (lo.value, lo.value + cnt.value) == (a, b), (n, world, r)


def test_two_rank_sharded_msm_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 37, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
