#!/usr/bin/env python
"""BASELINE.json configs[2]: 1280x720 mono stream (ZED2-shaped), 2000 features, 8 levels, lapping {0, 1000} — frames/s of
ORBextractor::operator() on one B200, device-resident (batches of 64 and 512 frames in HBM, CUDA events on the launching
stream) and through the host-facing batched call (64-frame groups, pinned buffers). Parity of 2 frames against the
oracle is asserted first. On 8 GPUs the 64-frame batch is 8 frames per GPU (frames are independent: weak scaling, see
bench.py --gpus N for the sharded run of configs[1])."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import ORBextractor, synth  # noqa: E402
from orb_slam3_fast_b200.synth import KP_DTYPE  # noqa: E402
from oracle import orbref  # noqa: E402  (checker only)


def main():
    W, H, NF, LAP = 1280, 720, 2000, (0, 1000)
    distinct = np.stack([synth.scene(H, W, 300 + s) for s in range(8)])
    out = {"workload": "configs[2]: 1280x720 mono, 2000 features, 8 levels, lapping {0,1000}", "rows": []}
    ref = orbref.Extractor(NF)
    st = torch.cuda.Stream()
    for B in (64, 512):
        ex = ORBextractor(NF, max_batch=B)
        cap = ex.capacity
        imgs = torch.from_numpy(np.ascontiguousarray(np.tile(distinct, (B // 8, 1, 1)))).cuda()
        d_kps = torch.empty((B, cap, 7), dtype=torch.int32, device="cuda")
        d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
        d_n = torch.empty(B, dtype=torch.int32, device="cuda")
        d_mono = torch.empty(B, dtype=torch.int32, device="cuda")
        d_status = torch.empty(B, dtype=torch.int32, device="cuda")

        def step():
            ex.extract_batch_device(imgs.data_ptr(), B, W, H, W, W * H, LAP, d_kps.data_ptr(), d_desc.data_ptr(), cap,
                                    d_n.data_ptr(), d_mono.data_ptr(), d_status.data_ptr(), st.cuda_stream)
        with torch.cuda.stream(st):
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            reps = 20 if B == 64 else 5
            e0.record(st)
            for _ in range(reps):
                step()
            e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert (d_status.cpu().numpy() == 0).all()
        n = d_n.cpu().numpy()
        kps = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28).copy().view(KP_DTYPE).reshape(B, cap)
        desc = d_desc.cpu().numpy()
        for f in (1, 6):
            m_r, k_r, d_r = ref(distinct[f], LAP)
            assert int(d_mono[f].item()) == m_r and np.array_equal(kps[f, :n[f]], k_r) and np.array_equal(desc[f, :n[f]], d_r)
        out["rows"].append({"frames_per_launch": B, "ms": ms, "frames_per_s_device_resident": B / (ms * 1e-3),
                            "keypoints_per_frame": float(n.mean())})
    # host-facing: 512 frames through orbx_extract_batch in 64-frame groups
    ex = ORBextractor(NF, max_batch=64)
    host = torch.from_numpy(np.ascontiguousarray(np.tile(distinct, (64, 1, 1)))).pin_memory().numpy()
    cap = ex.capacity
    keep = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in (512 * 4, 512 * 4, 512 * cap * 28, 512 * cap * 32)]
    outs = (keep[0].numpy().view(np.int32), keep[1].numpy().view(np.int32),
            keep[2].numpy().view(KP_DTYPE).reshape(512, cap), keep[3].numpy().reshape(512, cap, 32))
    for _ in range(2):
        ex.extract_batch(host, LAP, outs)
    t0 = time.perf_counter()
    for _ in range(3):
        ex.extract_batch(host, LAP, outs)
    dt = (time.perf_counter() - t0) / 3
    out["frames_per_s_end_to_end_512_frames"] = 512 / dt
    out["h2d_bytes_per_frame"] = W * H
    print(json.dumps(out))


if __name__ == "__main__":
    main()
