#!/usr/bin/env python
"""HBM throughput of the image front-end kernel (cv::cvtColor *2GRAY) on 1 x B200: 1024 frames of
752x480 BGR / BGRA per launch (1.1 / 1.5 GB read, 0.37 GB written: far beyond the 126 MB L2), CUDA events on the
launching stream. Usage: python tools/bench_frontend.py > gpurun_out/frontend.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from orb_slam3_fast_b200 import lib as _lib  # noqa: E402
from oracle import orbref  # noqa: E402  (checker only)


def main():
    L = _lib.lib()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    B, h, w = 1024, 480, 752
    st = torch.cuda.Stream()
    out = {"frames_per_launch": B, "size": [w, h], "peak_gbs": peak, "rows": []}
    for ch in (3, 4):
        src = torch.randint(0, 256, (B, h, w, ch), dtype=torch.uint8, device="cuda")
        dst = torch.empty((B, h, w), dtype=torch.uint8, device="cuda")

        def run():
            rc = L.orbx_cvt_gray_device(0, B, src.data_ptr(), w, h, w * ch, h * w * ch, ch, 0, dst.data_ptr(), w, h * w,
                                        st.cuda_stream)
            assert rc == 0
        with torch.cuda.stream(st):
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(st)
            for _ in range(20):
                run()
            e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        k = 5
        assert np.array_equal(dst[k].cpu().numpy(), orbref.cvt_gray(src[k].cpu().numpy(), False))
        gbs = B * h * w * (ch + 1) / (ms * 1e-3) / 1e9
        out["rows"].append({"channels": ch, "ms_per_launch": ms, "algorithmic_gbs": gbs, "frac_of_peak": gbs / peak,
                            "frames_per_s": B / (ms * 1e-3)})
    # ---- cv::remap rectification (src/System.cc:293-294): 1024 frames, one pair of maps shared by the batch ----
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    rng = np.random.default_rng(0)
    mapx = (xs + 3.7 * np.sin(ys / 50.0) + rng.uniform(-0.5, 0.5, (h, w)) - 4).astype(np.float32)
    mapy = (ys + 2.9 * np.cos(xs / 70.0) + rng.uniform(-0.5, 0.5, (h, w)) - 3).astype(np.float32)
    d_mx, d_my = torch.from_numpy(mapx).cuda(), torch.from_numpy(mapy).cuda()
    src = torch.randint(0, 256, (B, h, w), dtype=torch.uint8, device="cuda")
    dst = torch.empty((B, h, w), dtype=torch.uint8, device="cuda")

    def run_remap():
        rc = L.orbx_remap_linear_device(0, B, src.data_ptr(), w, h, w, h * w, d_mx.data_ptr(), d_my.data_ptr(), w, h,
                                        dst.data_ptr(), w, h * w, st.cuda_stream)
        assert rc == 0
    with torch.cuda.stream(st):
        for _ in range(3):
            run_remap()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(st)
        for _ in range(20):
            run_remap()
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    assert np.array_equal(dst[7].cpu().numpy(), orbref.remap_linear(src[7].cpu().numpy(), mapx, mapy))
    gbs = B * h * w * 2 / (ms * 1e-3) / 1e9   # 1 byte read + 1 byte written per pixel; the maps stay in L2
    out["remap_linear"] = {"ms_per_launch": ms, "algorithmic_gbs": gbs, "frac_of_peak": gbs / peak,
                           "frames_per_s": B / (ms * 1e-3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
