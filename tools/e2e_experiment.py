import os, sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth
P, W, H, G = 1024, 752, 480, int(os.environ.get("G", "64"))
pairs = [synth.stereo_pair(H, W, 1000 + s) for s in range(8)]
def pinned(shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    t = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True); pinned.keep.append(t)
    return t.numpy()[:n].view(dtype).reshape(shape)
pinned.keep = []
L = pinned((P, H, W), np.uint8); R = pinned((P, H, W), np.uint8)
for k in range(P):
    L[k] = pairs[k % 8][0]; R[k] = pairs[k % 8][1]
exl, exr = ORBextractor(1200, max_batch=G), ORBextractor(1200, max_batch=G)
mt = ORBmatcher()
outs = ORBmatcher.alloc_stereo_outputs(P, exl.capacity, empty=pinned)
mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
for _ in range(3): mt.StereoFramesBatch(exl, exr, L, R, mbf, mb, outs)
t0 = time.perf_counter()
for _ in range(5): mt.StereoFramesBatch(exl, exr, L, R, mbf, mb, outs)
dt = (time.perf_counter() - t0) / 5
print("G=%d skip_h2d=%s: %.2f ms per 2048 frames = %.0f fps" % (G, os.environ.get("ORBX_DEBUG_SKIP_H2D"), dt * 1e3, 2 * P / dt))
if os.environ.get("BG_H2D"):
    # the same calls without their own upload, while an unrelated stream keeps PCIe busy with H2D copies
    import threading
    stop = False
    src = torch.empty(46 * 1024 * 1024, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty_like(src, device="cuda")
    cs = torch.cuda.Stream()
    n_copies = [0]
    def pump():
        with torch.cuda.stream(cs):
            while not stop:
                for _ in range(4):
                    dst.copy_(src, non_blocking=True); n_copies[0] += 1
                cs.synchronize()
    th = threading.Thread(target=pump); th.start()
    time.sleep(0.2)
    c0 = n_copies[0]; t0 = time.perf_counter()
    for _ in range(5): mt.StereoFramesBatch(exl, exr, L, R, mbf, mb, outs)
    dt = (time.perf_counter() - t0) / 5; c1 = n_copies[0]
    stop = True; th.join()
    print("with background H2D (%.1f GB/s): %.2f ms per 2048 frames" % ((c1 - c0) * 46 * 1.048576e6 / (5 * dt) / 1e9, dt * 1e3))
