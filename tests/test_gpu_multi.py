"""Several GPUs behind ONE C-ABI call (msm_b200_multi_*): the range-sharded result must equal the single-GPU
result and the CPU port, for every device count the box offers.  No torch in the product path here: the
library drives the devices from its own host threads and gathers the partials with NCCL or peer copies.
Plus the torchrun / NCCL harness (msm_zprize_b200/dist.py) on 2 ranks when 2 GPUs are visible."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mz():
    import msm_zprize_b200 as m
    return m


def _device_count():
    out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
    return sum(1 for l in out.splitlines() if l.startswith("GPU "))


def _device_sets():
    n = _device_count()
    sets = [[0]]
    for k in (2, 4, 8):
        if n >= k:
            sets.append(list(range(k)))
    if n >= 3:
        sets.append([2, 0, 1])  # order = shard order, not device numbering
    return sets


@pytest.mark.parametrize("name", ["bls12-377", "ed-on-bls12-377"])
def test_multi_equals_single_and_port(mz, name):
    from oracle.port import Port
    n = (1 << 16) + 77  # ragged: the last shard is shorter
    threads = os.cpu_count() or 1
    port = Port(name)
    pts = port.random_points(n, 0xB216, threads)
    sc = port.random_scalars(n, 0x5C16, threads)
    want = port.msm(sc, port.prepare_points(pts, n, threads), n, threads)[:3]
    with mz.MsmEngine(name) as eng:
        single = eng.msm(sc, pts, n)
    assert (single.x, single.y, single.is_zero) == want
    for devices in _device_sets():
        for gather in (None, "peer"):
            if gather:
                os.environ["MSM_B200_GATHER"] = gather
            else:
                os.environ.pop("MSM_B200_GATHER", None)
            try:
                with mz.MultiMsmEngine(name, devices) as m:
                    if gather and len(devices) > 1:
                        assert "peer" in m.gather_kind
                    m.set_bases(pts, n)
                    r = m.run(sc, n)
                    assert (r.x, r.y, r.is_zero) == want, (devices, m.gather_kind)
                    assert r.timing["kernel_launches"] > 0
                    # a prefix of the resident set: trailing shards may be empty
                    k = n // 3
                    rk = m.run(sc, k)
                    with mz.MsmEngine(name) as eng:
                        sk = eng.msm(sc, pts, k)
                    assert (rk.x, rk.y, rk.is_zero) == (sk.x, sk.y, sk.is_zero)
                    # one-shot (points and scalars from the host in the same call)
                    r1 = m.msm(sc, pts, n)
                    assert (r1.x, r1.y, r1.is_zero) == want
                    assert len(m.last_timings()) == len(devices)
                    # fewer points than devices: most shards are empty
                    r5 = m.msm(sc[:5 * 32], pts[:5 * 2 * port.nbytes], 5)
                    with mz.MsmEngine(name) as eng:
                        s5 = eng.msm(sc[:5 * 32], pts[:5 * 2 * port.nbytes], 5)
                    assert (r5.x, r5.y, r5.is_zero) == (s5.x, s5.y, s5.is_zero)
                    r0 = m.msm(b"", b"", 0)
                    assert r0.is_zero
            finally:
                os.environ.pop("MSM_B200_GATHER", None)


def test_multi_sharded_device_inputs(mz):
    """Shards generated on their own devices (msm_b200_random_points_at) == the same seeded set on one GPU."""
    name, n = "bls12-377", 1 << 15
    with mz.MsmEngine(name) as eng:
        pb = eng.point_bytes(mz.LAYOUT_LE_BYTES)
        d_pts, d_sc = eng.dev_alloc(n * pb), eng.dev_alloc(n * 32)
        eng.random_points_device(d_pts, n, 0xB215)
        eng.random_scalars_device(d_sc, n, 0x5C15)
        eng.set_bases_device(d_pts, n)
        want = eng.run(d_sc, n, on_device=True)
    for devices in _device_sets():
        with mz.MultiMsmEngine(name, devices) as m:
            G = len(devices)
            per = -(-n // G)
            pp, ss, cnt = [], [], []
            for g, e in enumerate(m.shards):
                lo = min(n, per * g)
                k = min(n, lo + per) - lo
                p, s = e.dev_alloc(max(k, 1) * pb), e.dev_alloc(max(k, 1) * 32)
                if k:
                    e.random_points_device(p, k, 0xB215, first=lo)
                    e.random_scalars_device(s, k, 0x5C15, first=lo)
                pp.append(p), ss.append(s), cnt.append(k)
            m.set_bases_sharded(pp, cnt)
            r = m.run_sharded(ss)
            assert (r.x, r.y, r.is_zero) == (want.x, want.y, want.is_zero), devices


def test_multi_contexts_share_bases(mz):
    """Two multi contexts over ONE resident point set (msm_b200_multi_share_bases), driven from two host threads."""
    import threading
    from oracle.port import Port
    name, n = "pallas", (1 << 14) + 9
    port = Port(name)
    pts = port.random_points(n, 81, 4)
    scs = [port.random_scalars(n, 90 + i, 4) for i in range(4)]
    prep = port.prepare_points(pts, n, 4)
    want = [port.msm(s, prep, n, 4)[:3] for s in scs]
    for devices in _device_sets()[:2]:
        with mz.MultiMsmEngine(name, devices) as a, mz.MultiMsmEngine(name, devices) as b:
            a.set_bases(pts, n)
            b.share_bases(a)
            got = [None] * 4

            def work(m, idx):
                for i in idx:
                    r = m.run(scs[i], n)
                    got[i] = (r.x, r.y, r.is_zero)

            th = [threading.Thread(target=work, args=(a, (0, 2))), threading.Thread(target=work, args=(b, (1, 3)))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert got == want, devices
            b.close()  # the borrower goes first


def test_pipeline_api_submit_wait(mz):
    """msm_b200_pipeline_*: plain submit / wait calls, the lanes' threads live inside the library."""
    from oracle.port import Port
    name, n = "bls12-377", (1 << 14) + 3
    port = Port(name)
    pts = port.random_points(n, 101, 4)
    scs = [port.random_scalars(n, 110 + i, 4) for i in range(7)]
    prep = port.prepare_points(pts, n, 4)
    want = [port.msm(s, prep, n, 4)[:3] for s in scs]
    for devices in _device_sets()[:2]:
        with mz.MsmPipeline(name, devices, depth=3) as p:
            p.set_bases(pts, n)
            tickets = [p.submit(s, n) for s in scs]  # 7 MSMs over 3 lanes, none awaited in between
            got = [p.wait(t) for t in tickets]
            assert [(r.x, r.y, r.is_zero) for r in got] == want, devices
            assert all(r.timing["shared_buckets"] == 1 for r in got)
            # new bases while nothing is in flight; a short scalar buffer is refused before the call
            pts2 = port.random_points(n, 102, 4)
            p.set_bases(pts2, n)
            r = p.wait(p.submit(scs[0], n))
            assert (r.x, r.y, r.is_zero) == port.msm(scs[0], port.prepare_points(pts2, n, 4), n, 4)[:3]
            with pytest.raises(mz.MsmError):
                p.submit(bytes(10), n)
    with pytest.raises(mz.MsmError):
        mz.MsmPipeline(name, [0], depth=0)
    with mz.MsmPipeline(name, [0], depth=2) as p:
        t = p.submit(scs[0], n)  # no bases yet: the job fails with a code, delivered by wait()
        with pytest.raises(mz.MsmError):
            p.wait(t)
        p.set_bases(pts, n)  # and the pipeline is usable afterwards
        r = p.wait(p.submit(scs[1], n))
        assert (r.x, r.y, r.is_zero) == want[1]
        p.submit(scs[2], n)  # left in flight: destroy waits for it


def test_multi_errors_are_codes(mz):
    with pytest.raises(mz.MsmError):
        mz.MultiMsmEngine("bls12-377", [0, 0])  # duplicate device
    with pytest.raises(mz.MsmError):
        mz.MultiMsmEngine("bls12-377", [999])
    with mz.MultiMsmEngine("pallas", [0]) as m:
        with pytest.raises(mz.MsmError):
            m.run(bytes(32 * 4), 4)  # no bases yet
        with pytest.raises(mz.MsmError):
            m.set_bases(bytes(10), 4)  # short buffer: refused, not read past its end


_TORCHRUN_BODY = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from msm_zprize_b200.dist import ShardedMsm, shard_range
from oracle.port import Port
n = (1 << 16) + 5
port = Port("bls12-377")
pts = port.random_points(n, 0xB216, 4)
sc = port.random_scalars(n, 0x5C16, 4)
lo, hi = shard_range(n, rank, world)
sm = ShardedMsm("bls12-377", device=local)
sm.set_bases(pts[lo * 96:hi * 96], hi - lo)
res = sm.msm(sc[lo * 32:hi * 32], hi - lo)
if rank == 0:
    want = port.msm(sc, port.prepare_points(pts, n, 8), n, 8)[:3]
    assert (res.x, res.y, res.is_zero) == want, "sharded NCCL result differs from the CPU port"
    print("SHARDED_OK", world)
dist.barrier()
dist.destroy_process_group()
"""


def test_torchrun_nccl_sharded_equals_port(tmp_path):
    n = _device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 4 if n >= 4 else 2
    script = tmp_path / "sharded.py"
    script.write_text(_TORCHRUN_BODY.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and f"SHARDED_OK {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
