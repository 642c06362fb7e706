"""Multi-rank host logic on CPU: world_size 2 (and 3) over gloo. Each rank extracts its shard of a batch — with the CPU
oracle standing in for the device path, since there is no GPU here — and the product's FrameSharder gathers the result
slabs; every rank must end up with exactly what a single process computes for the whole batch, in frame order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from orb_slam3_fast_b200 import sharding  # noqa: E402
from orb_slam3_fast_b200.lib import KP_DTYPE  # noqa: E402


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 7, 8, 63, 64, 65, 1000):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == sharding.shard_sizes(n, world)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _frames(n):
    from orb_slam3_fast_b200 import synth
    return np.stack([synth.scene(241, 260, seed=50 + i) for i in range(n)])


def _extract_with_oracle(imgs, cap):
    from oracle import orbref
    ex = orbref.Extractor(300)
    n_out = np.zeros(len(imgs), np.int32)
    mono = np.zeros(len(imgs), np.int32)
    kps = np.zeros((len(imgs), cap), KP_DTYPE)
    desc = np.zeros((len(imgs), cap, 32), np.uint8)
    for i, im in enumerate(imgs):
        m, k, d = ex(im, (0, 100))
        n_out[i], mono[i] = len(k), m
        kps[i, :len(k)] = k
        desc[i, :len(k)] = d
    return n_out, mono, kps, desc


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap = 300 + 16 * 8
        sh = sharding.FrameSharder()
        a, b = sh.my_range(n_frames)
        imgs = _frames(n_frames)
        local = _extract_with_oracle(imgs[a:b], cap)
        g = sh.gather_extract_results(n_frames, *local)
        ref = _extract_with_oracle(imgs, cap)
        ok = all(np.array_equal(x, y) for x, y in zip(g, ref))
        counts = sh.gather_counts(torch.from_numpy(np.stack([local[0][:1], local[1][:1]])))
        ok = ok and tuple(counts.shape) == (world, 2, 1)
        q.put((rank, bool(ok), int(g[0].sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (2, 4), (3, 7)])
def test_sharded_extract_gathers_to_the_serial_result(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert len({s for _, _, s in res}) == 1 and res[0][2] > 0


def _knn_worker(rank, world, port, nq, nt, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from orb_slam3_fast_b200 import synth
        from oracle import orbref
        qd, td = synth.descriptors(nq, 7, prototypes=16), synth.descriptors(nt, 8, prototypes=16)  # tie heavy
        sh = sharding.FrameSharder()
        a, b = sh.my_range(nq)
        local = orbref.knn2(qd[a:b], td)            # the CPU oracle stands in for the per-rank GPU call
        g = sh.gather_knn2(nq, *local)
        ref = orbref.knn2(qd, td)
        q.put((rank, bool(all(np.array_equal(x, y) for x, y in zip(g, ref)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nq,nt", [(2, 101, 300), (3, 50, 64)])
def test_sharded_knn2_gathers_to_the_serial_result(world, nq, nt):
    """configs[4] on several GPUs: queries sharded, train set replicated, one gather of 16 B per query."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_worker, args=(r, world, port, nq, nt, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=90) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
