"""bench.py --config 2 / --config 4: the two BASELINE.json configurations that are not stereo-pair workloads, as JSON
lines in bench.py's format (rank 0 emits). Launched like bench.py itself (python for N = 1, torchrun for N > 1).

configs[2]: 1280x720 mono stream (ZED2-shaped), 2000 features, 8 levels, lapping {0, 1000}, ONE batch of 64 frames sharded
            over the N GPUs (64 / N frames per GPU: total work fixed -> "scaling": "strong"); the result slabs are
            gathered to every rank with one all_gather (FrameSharder.gather_extract_results' device form).
configs[4]: N x N brute-force 256-bit Hamming 2-NN for N = 1k ... 100k, queries sharded over the GPUs, train set
            replicated, one all_gather of 16 B per query; value = pairs/s at 100k x 100k.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _pctl(xs, q):
    xs = sorted(xs)
    k = (len(xs) - 1) * q
    lo, hi = int(np.floor(k)), int(np.ceil(k))
    return xs[lo] + (xs[hi] - xs[lo]) * (k - lo)


def run_config2(ctx, args, metric):
    torch, dist = ctx.torch, ctx.dist
    from orb_slam3_fast_b200 import ORBextractor, sharding, synth
    from orb_slam3_fast_b200.lib import KP_DTYPE
    W, H, NF, LAP, TOTAL = 1280, 720, 2000, (0, 1000), 64
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    a, b = sharding.shard_range(TOTAL, world, rank)
    B = b - a
    imgs_h = np.stack([synth.scene(H, W, 300 + s) for s in range(a, b)])
    ex = ORBextractor(NF, device=ctx.local, max_batch=B)
    cap = ex.capacity
    n_rot = 2
    imgs = [torch.from_numpy(np.ascontiguousarray(np.roll(imgs_h, k, axis=0))).to(dev) for k in range(n_rot)]
    pin = [ctx.pinned(imgs_h.shape, np.uint8) for _ in range(n_rot)]
    for k in range(n_rot):
        pin[k][...] = np.roll(imgs_h, k, axis=0)
    d_kps = torch.empty((B, cap, 7), dtype=torch.int32, device=dev)
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.empty((3, B), dtype=torch.int32, device=dev)  # n, mono, status
    slab = torch.empty((B, cap * 15 + 2), dtype=torch.int32, device=dev)  # count + mono + 28-B keypoints + 32-B descriptors
    gathered = torch.empty((world * B, cap * 15 + 2), dtype=torch.int32, device=dev) if world > 1 else None
    st = ctx.stream.cuda_stream

    def step(k):
        ex.extract_batch_device(imgs[k % n_rot].data_ptr(), B, W, H, W, W * H, LAP, d_kps.data_ptr(), d_desc.data_ptr(),
                                cap, d_n[0].data_ptr(), d_n[1].data_ptr(), d_n[2].data_ptr(), st)
        if world > 1:  # the result gather: one fixed-size slab per frame to every rank
            slab[:, 0] = d_n[0]
            slab[:, 1] = d_n[1]
            slab[:, 2:2 + cap * 7] = d_kps.view(B, cap * 7)
            slab[:, 2 + cap * 7:] = d_desc.view(B, cap * 32).view(torch.int32)
            dist.all_gather_into_tensor(gathered, slab)

    # parity gate: this rank's first two frames against the oracle (rank 0 decides; every rank checks its own)
    from oracle import orbref
    step(0)
    torch.cuda.synchronize()
    n = d_n[0].cpu().numpy()
    kps = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28).copy().view(KP_DTYPE).reshape(B, cap)
    desc = d_desc.cpu().numpy()
    ref = orbref.Extractor(NF)
    for f in range(min(2, B)):
        m_r, k_r, d_r = ref(imgs_h[f], LAP)
        if not (int(d_n[1][f].item()) == m_r and np.array_equal(kps[f, :n[f]], k_r) and np.array_equal(desc[f, :n[f]], d_r)):
            raise SystemExit("bench.py --config 2: parity against the oracle FAILED on rank %d frame %d" % (rank, f))
    reps = args.reps * 8  # a 64-frame batch takes ~1 ms: a step is `reps` batches
    for k in range(max(args.warmup, 3) * reps):
        step(k)
    ctx.barrier()
    nb = args.steps * reps
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(nb + 1)]
    evs[0].record()
    for k in range(nb):
        step(k)
        evs[k + 1].record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(evs[0].elapsed_time(evs[nb]))
    per = [evs[k].elapsed_time(evs[k + 1]) for k in range(nb)]
    value = TOTAL * nb / (ms_total * 1e-3)
    # end to end: the host-facing batched call on this rank's shard
    outs = (ctx.pinned((B,), np.int32), ctx.pinned((B,), np.int32), ctx.pinned((B, cap), KP_DTYPE),
            ctx.pinned((B, cap, 32), np.uint8))
    ex2 = ORBextractor(NF, device=ctx.local, max_batch=max(1, min(16, B)))
    for k in range(3):
        ex2.extract_batch(pin[k % n_rot], LAP, outs)
    ctx.barrier()
    t0 = time.perf_counter()
    for k in range(nb):
        ex2.extract_batch(pin[k % n_rot], LAP, outs)
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[2]: 1280x720 mono stream (ZED2-shaped), 2000 features, 8 levels, lapping "
                                   "{0,1000}, batched 64 frames sharded over the GPUs (%d per GPU), result slabs "
                                   "all-gathered" % B,
                       "batches_per_step": reps, "frames_per_batch_total": TOTAL,
                       "ms_per_batch": {"p10": _pctl(per, 0.1), "p50": _pctl(per, 0.5), "p90": _pctl(per, 0.9), "n": nb},
                       "keypoints_per_frame": float(n.mean()),
                       "parity": "bit-exact vs oracle on 2 frames per rank (mono index, keypoints, descriptors)",
                       "parallelism": "frames sharded over GPUs; one all_gather of the result slabs per batch"},
            "gpu_launches": int(ex._L.orbx_kernel_launches(ex._h)) * nb,
            "e2e": {"value": TOTAL * nb / dt, "unit": "frames/s", "h2d_bytes_per_step": B * W * H * reps,
                    "d2h_bytes_per_step": B * (cap * 60 + 12) * reps,
                    "api": "orbx_extract_batch (host buffers, pinned) on each rank's shard"}}
    return line


def run_config4(ctx, args, metric):
    torch, dist = ctx.torch, ctx.dist
    from orb_slam3_fast_b200 import ORBmatcher, sharding, synth
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    mt = ORBmatcher(device=ctx.local)
    rows = []
    st = ctx.stream.cuda_stream
    for n in (1000, 3000, 10000, 30000, 100000):
        q, t = synth.descriptors(n, 21), synth.descriptors(n, 22)
        a, b = sharding.shard_range(n, world, rank)
        m = b - a
        mmax = max(sharding.shard_sizes(n, world))
        dq, dt_ = torch.from_numpy(q[a:b]).to(dev), torch.from_numpy(t).to(dev)
        o = torch.full((4, mmax), -1, dtype=torch.int32, device=dev)
        g = torch.empty((world * 4, mmax), dtype=torch.int32, device=dev) if world > 1 else None

        def step():
            mt.knnMatch2_device(dq.data_ptr(), m, dt_.data_ptr(), n, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                o[3].data_ptr(), st)
            if world > 1:
                dist.all_gather_into_tensor(g, o)
        step()
        torch.cuda.synchronize()
        res = (g.view(world, 4, mmax) if world > 1 else o.view(1, 4, mmax)).cpu().numpy()
        sizes = sharding.shard_sizes(n, world)
        idx1 = np.concatenate([res[r, 0, :sizes[r]] for r in range(world)])
        idx2 = np.concatenate([res[r, 2, :sizes[r]] for r in range(world)])
        rr = np.random.default_rng(n).integers(0, n, 12)
        D = np.bitwise_count(q[rr].view(np.uint64)[:, None, :] ^ t.view(np.uint64)[None, :, :]).sum(axis=2)
        order = np.lexsort((np.broadcast_to(np.arange(n), D.shape), D), axis=1)[:, :2]
        if not (np.array_equal(idx1[rr], order[:, 0]) and np.array_equal(idx2[rr], order[:, 1])):
            raise SystemExit("bench.py --config 4: knn2 parity FAILED at n = %d" % n)
        reps = max(5, min(200, int(2e10 / (n * n / world + 1))))
        for _ in range(3):
            step()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            step()
        e1.record()
        ctx.barrier()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / reps
        rows.append({"n": n, "ms": ms, "gpairs_per_s": n * n / ms / 1e6, "reps": reps})
    big = rows[-1]
    line = {"metric": "pairs/sec brute-force 256-bit Hamming 2-NN (BASELINE.json configs[4]); " + metric,
            "value": big["gpairs_per_s"] * 1e9, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": 3, "ms_per_step": big["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 (s8 x s8 -> s32 on the tensor cores above the switch-over, POPC below)", "data": "synthetic",
            "config": {"workload": "configs[4]: N x N Hamming 2-NN sweep, queries sharded over the GPUs, train set "
                                   "replicated, one all_gather of 16 B per query", "rows": rows,
                       "parity": "12 sampled rows per size vs numpy (distance, index) order",
                       "timer": "CUDA events on the launching stream, max over ranks; includes the result gather"},
            "gpu_launches": 3, "e2e": None}
    return line
