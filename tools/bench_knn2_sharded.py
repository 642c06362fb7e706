#!/usr/bin/env python
"""BASELINE.json configs[4] on N GPUs of one box: brute-force 256-bit Hamming 2-NN, queries sharded over the ranks,
train set replicated, one all_gather of 16 B per query over NCCL. Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
      tools/bench_knn2_sharded.py          (or plain `python tools/bench_knn2_sharded.py` for N = 1)
Rank 0 prints one JSON object. Time = max over ranks of (device-resident knn2 on the shard + the gather), CUDA-synced
wall clock; parity of the gathered result against numpy on sampled rows is asserted first."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import ORBmatcher, sharding, synth  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sh = sharding.FrameSharder(dev)
    mt = ORBmatcher(device=local)
    rows = []
    for n in (10000, 30000, 100000):
        q, t = synth.descriptors(n, 21), synth.descriptors(n, 22)
        a, b = sh.my_range(n)
        dq, dt = torch.from_numpy(q[a:b]).to(dev), torch.from_numpy(t).to(dev)
        m = b - a
        o = [torch.empty(max(m, 1), dtype=torch.int32, device=dev) for _ in range(4)]

        def step():
            mt.knnMatch2_device(dq.data_ptr(), m, dt.data_ptr(), n, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                o[3].data_ptr())
            torch.cuda.synchronize()  # the matcher launches on its own stream
            return sh.gather_knn2(n, o[0][:m], o[1][:m], o[2][:m], o[3][:m])
        g = step()
        rr = np.random.default_rng(n).integers(0, n, 12)
        D = np.bitwise_count(q[rr].view(np.uint64)[:, None, :] ^ t.view(np.uint64)[None, :, :]).sum(axis=2)
        order = np.lexsort((np.broadcast_to(np.arange(n), D.shape), D), axis=1)[:, :2]
        assert np.array_equal(g[0][rr], order[:, 0]) and np.array_equal(g[2][rr], order[:, 1])
        reps = 5
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        torch.cuda.synchronize()
        dt_s = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt_s, op=dist.ReduceOp.MAX)
        ms = float(dt_s.item()) * 1e3
        rows.append({"n": n, "ms": ms, "gpairs_per_s": n * n / ms / 1e6})
    if rank == 0:
        print(json.dumps({"workload": "configs[4]: N x N Hamming 2-NN, queries sharded, train replicated", "n_gpus": world,
                          "rows": rows, "includes": "device knn2 on the shard + all_gather of the results to every rank"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
