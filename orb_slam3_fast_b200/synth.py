"""Seeded synthetic inputs for the ORB front-end (SURVEY.md §8d "Synthetic images"). numpy only, deterministic.

There is no network for EuRoC / TUM-VI, so tests and bench.py use these generators; `data: "synthetic"` in the
bench line refers to them.
"""
import numpy as np

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


def _bilinear_up(a, h, w):
    gh, gw = a.shape
    ys = np.linspace(0, gh - 1, h)
    xs = np.linspace(0, gw - 1, w)
    y0 = np.minimum(ys.astype(np.int64), gh - 2)
    x0 = np.minimum(xs.astype(np.int64), gw - 2)
    fy = (ys - y0)[:, None]
    fx = (xs - x0)[None, :]
    a00 = a[y0][:, x0]
    a01 = a[y0][:, x0 + 1]
    a10 = a[y0 + 1][:, x0]
    a11 = a[y0 + 1][:, x0 + 1]
    return (a00 * (1 - fx) + a01 * fx) * (1 - fy) + (a10 * (1 - fx) + a11 * fx) * fy


def scene(h, w, seed=0, n_rect=150):
    """5-octave value noise + random rectangles, normalised to 0..255: ~7-10 k FAST candidates at 640x480."""
    rng = np.random.default_rng(seed)
    img = np.zeros((h, w), np.float64)
    for cell, amp in ((4, 60), (8, 40), (16, 30), (32, 20), (64, 10)):
        g = rng.random((h // cell + 3, w // cell + 3))
        img += amp * _bilinear_up(g, h, w)
    for _ in range(n_rect):
        rw, rh = rng.integers(8, 80, 2)
        x, y = rng.integers(0, w - 8), rng.integers(0, h - 8)
        img[y:y + rh, x:x + rw] += rng.uniform(-60, 60)
    img -= img.min()
    img *= 255.0 / max(img.max(), 1e-9)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def _box_blur(img, k):
    out = img.astype(np.float64)
    for axis in (0, 1):
        acc = np.zeros_like(out)
        for d in range(-k, k + 1):
            acc += np.roll(out, d, axis=axis)
        out = acc / (2 * k + 1)
    return out


def noise_blur(h, w, seed=0):
    rng = np.random.default_rng(seed)
    img = _box_blur(rng.integers(0, 256, (h, w)).astype(np.float64), 1)
    img = (img - img.min()) * (255.0 / (img.max() - img.min()))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def uniform_noise(h, w, seed=0):
    return np.random.default_rng(seed).integers(0, 256, (h, w), dtype=np.uint8)


def constant(h, w, value=127):
    return np.full((h, w), value, np.uint8)


def low_contrast(h, w, seed=0, amplitude=12):
    """Amplitude <= 15 around mid-grey: no pixel passes iniThFAST=20, so every cell takes the minThFAST retry."""
    img = scene(h, w, seed).astype(np.float64)
    return np.clip(np.rint(128 + (img - 128) * (amplitude / 128.0)), 0, 255).astype(np.uint8)


KINDS = {"scene": scene, "noise_blur": noise_blur, "uniform_noise": uniform_noise}


def make(kind, h, w, seed=0):
    return KINDS[kind](h, w, seed)


def stereo_pair(h, w, seed=0, d_min=2, d_max=60, kind="scene"):
    """Left image + right image = left shifted left by a per-row-band constant disparity (so true matches exist).

    Returns (left, right, disparity_per_row)."""
    rng = np.random.default_rng(seed + 7919)
    wide = make(kind, h, w + d_max + 1, seed)
    left = np.ascontiguousarray(wide[:, :w])
    disp = np.empty(h, np.int64)
    y = 0
    while y < h:
        band = int(rng.integers(24, 64))
        disp[y:y + band] = int(rng.integers(d_min, d_max + 1))
        y += band
    right = np.empty_like(left)
    cols = np.arange(w)
    for r in range(h):
        right[r] = wide[r, cols + disp[r]]
    return left, right, disp


def descriptors(n, seed=0, prototypes=0):
    """i.i.d. uniform 256-bit descriptors; prototypes > 0 draws rows from that many prototypes (tie-heavy)."""
    rng = np.random.default_rng(seed)
    if prototypes:
        protos = rng.integers(0, 256, (prototypes, 32), dtype=np.uint8)
        return np.ascontiguousarray(protos[rng.integers(0, prototypes, n)])
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def flip_bits(desc, k, rng):
    """Copy of desc (n,32) with k[i] random bit flips in row i."""
    out = desc.copy()
    for i in range(len(out)):
        bits = rng.choice(256, size=int(k[i]), replace=False)
        for b in bits:
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def local_map(kps, desc, m, w, h, n_levels, seed=0, bf=47.9):
    """Config 4: m synthetic local-map points against one frame. Returns a dict of SoA arrays in
    the orbx_mappoints layout. Half of the points are anchored on real keypoints (projection jittered, descriptor =
    the keypoint's with 0..80 flipped bits), the rest are uniform over the image."""
    rng = np.random.default_rng(seed + 104729)
    n = len(kps)
    src = rng.integers(0, n, m)
    anchored = rng.random(m) < 0.5
    px = np.where(anchored, kps["x"][src] + rng.normal(0, 1.5, m), rng.uniform(0, w, m)).astype(np.float32)
    py = np.where(anchored, kps["y"][src] + rng.normal(0, 1.5, m), rng.uniform(0, h, m)).astype(np.float32)
    level = np.where(anchored, np.clip(kps["octave"][src] + rng.integers(0, 2, m), 0, n_levels - 1),
                     rng.integers(0, n_levels, m)).astype(np.int32)
    depth = rng.uniform(0.5, 20.0, m).astype(np.float32)
    d = flip_bits(desc[src], rng.integers(0, 81, m), rng)
    return dict(track_in_view=(rng.random(m) < 0.95).astype(np.uint8), proj_x=px, proj_y=py,
                proj_xr=(px - np.float32(bf) / depth).astype(np.float32), level=level,
                view_cos=rng.uniform(0.5, 1.0, m).astype(np.float32), depth=depth,
                has_obs=(rng.random(m) < 0.9).astype(np.uint8), desc=d)


def camera_pose(seed=0):
    """A seeded world-to-camera pose (Rcw, tcw) and the camera centre Ow = -Rcw^T tcw, float32 (Frame::SetPose /
    UpdatePoseMatrices, src/Frame.cc:484-518)."""
    rng = np.random.default_rng(seed + 7919)
    a = rng.uniform(-0.6, 0.6, 3)
    cx_, sx = np.cos(a[0]), np.sin(a[0])
    cy_, sy = np.cos(a[1]), np.sin(a[1])
    cz_, sz = np.cos(a[2]), np.sin(a[2])
    R = (np.array([[cz_, -sz, 0], [sz, cz_, 0], [0, 0, 1]]) @ np.array([[cy_, 0, sy], [0, 1, 0], [-sy, 0, cy_]]) @
         np.array([[1, 0, 0], [0, cx_, -sx], [0, sx, cx_]]))
    t = rng.uniform(-2.0, 2.0, 3)
    R32, t32 = R.astype(np.float32), t.astype(np.float32)
    Ow = (-(R32.astype(np.float64).T @ t32.astype(np.float64))).astype(np.float32)
    return R32, t32, Ow


def frustum(w, h, seed=0, fx=435.2, fy=435.2, cx=None, cy=None, bf=47.9, scale_factor=1.2, n_levels=8):
    """One orbx_frustum record (views.FRUSTUM_DTYPE) for an undistorted w x h pinhole camera at camera_pose(seed)."""
    from .views import FRUSTUM_DTYPE
    R, t, Ow = camera_pose(seed)
    fr = np.zeros((), FRUSTUM_DTYPE)
    fr["Rcw"], fr["tcw"], fr["Ow"] = R.reshape(9), t, Ow
    fr["fx"], fr["fy"] = fx, fy
    fr["cx"] = (w - 1) / 2.0 if cx is None else cx
    fr["cy"] = (h - 1) / 2.0 if cy is None else cy
    fr["mbf"] = bf
    fr["min_x"], fr["max_x"], fr["min_y"], fr["max_y"] = 0.0, w, 0.0, h
    fr["log_scale_factor"] = np.log(np.float32(scale_factor))   # float32 log, as logf(mfScaleFactor) (src/Frame.cc:189)
    fr["n_levels"] = n_levels
    return fr


def local_map_world(kps, desc, m, fr, seed=0):
    """Config 4 with the projection on the device: m world points of a synthetic local map seen from frustum `fr`
    (SURVEY.md §8d): ~half are anchored on real keypoints of the frame (projection jittered by ~1.5 px, predicted level =
    the keypoint's octave or one above, descriptor = the keypoint's with 0..80 flipped bits), the rest project uniformly
    over (and slightly outside) the image; a few are behind the camera, out of their distance range or seen too
    obliquely. Returns a dict in the orbx_local_map layout."""
    rng = np.random.default_rng(seed + 611953)
    n = len(kps)
    w, h = float(fr["max_x"]), float(fr["max_y"])
    n_levels = int(fr["n_levels"])
    src = rng.integers(0, max(n, 1), m)
    anchored = (rng.random(m) < 0.5) & (n > 0)
    kx = kps["x"][src] if n else np.zeros(m)
    ky = kps["y"][src] if n else np.zeros(m)
    ko = kps["octave"][src] if n else np.zeros(m, np.int64)
    px = np.where(anchored, kx + rng.normal(0, 1.5, m), rng.uniform(-20, w + 20, m))
    py = np.where(anchored, ky + rng.normal(0, 1.5, m), rng.uniform(-20, h + 20, m))
    level = np.where(anchored, np.clip(ko + rng.integers(0, 2, m), 0, n_levels - 1), rng.integers(0, n_levels, m))
    z = rng.uniform(0.5, 20.0, m)
    z = np.where(rng.random(m) < 0.02, -z, z)                                    # behind the camera
    R = fr["Rcw"].reshape(3, 3).astype(np.float64)
    t = fr["tcw"].astype(np.float64)
    Pc = np.stack([(px - float(fr["cx"])) / float(fr["fx"]) * z, (py - float(fr["cy"])) / float(fr["fy"]) * z, z], 1)
    P = (Pc - t) @ R                                                             # Rwc (Pc - tcw)
    pos = P.astype(np.float32)
    PO = pos.astype(np.float64) - fr["Ow"].astype(np.float64)
    dist = np.linalg.norm(PO, axis=1)
    # mean viewing direction: the ray to the camera centre tilted by up to ~70 degrees (cos limit 0.5 = 60 degrees)
    tilt = rng.uniform(0, 1.2, m)
    axis = np.cross(PO, rng.normal(size=(m, 3)))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True) + 1e-12
    d0 = PO / (dist[:, None] + 1e-12)
    nrm = d0 * np.cos(tilt)[:, None] + np.cross(axis, d0) * np.sin(tilt)[:, None]
    # PredictScale = ceil(log(max_dist / dist) / log(scale)) -> aim at the middle of the level's interval
    sf = float(np.exp(fr["log_scale_factor"]))
    max_dist = dist * sf ** (level - 0.5)
    min_dist = max_dist / sf ** (n_levels - 1)
    bad_range = rng.random(m) < 0.05
    max_dist = np.where(bad_range, dist * 0.5, max_dist)                         # dist > 1.2 * max_dist
    flips = rng.integers(0, 81, m)
    d = flip_bits(desc[src], flips, rng) if n else rng.integers(0, 256, (m, 32), dtype=np.uint8)
    return dict(pos=pos, normal=nrm.astype(np.float32), min_dist=min_dist.astype(np.float32),
                max_dist=max_dist.astype(np.float32), skip=(rng.random(m) < 0.08).astype(np.uint8),
                has_obs=(rng.random(m) < 0.9).astype(np.uint8), desc=d)


def fisheye_case(n_left=700, n_right=650, m=3000, w=640, h=480, n_levels=8, seed=0):
    """A two-camera (Nleft != -1) frame and local-map points for SearchByProjection's fisheye branches: random left /
    right keypoints, ~40 % of them associated left <-> right (mvLeftToRightMatch / mvRightToLeftMatch, a partial
    bijection), points anchored on left and / or right keypoints with jittered projections and 0..70 flipped bits."""
    rng = np.random.default_rng(seed + 271828)
    def kps(n):
        k = np.zeros(n, KP_DTYPE)
        k["x"], k["y"] = rng.uniform(0, w, n).astype(np.float32), rng.uniform(0, h, n).astype(np.float32)
        k["octave"] = rng.integers(0, n_levels, n)
        k["size"], k["angle"], k["class_id"] = 31, rng.uniform(0, 360, n).astype(np.float32), -1
        return k
    kl, kr = kps(n_left), kps(n_right)
    dl = descriptors(n_left, seed + 1)
    dr = descriptors(n_right, seed + 2)
    l2r = np.full(n_left, -1, np.int32)
    r2l = np.full(n_right, -1, np.int32)
    npair = int(0.4 * min(n_left, n_right))
    pl, pr = rng.choice(n_left, npair, replace=False), rng.choice(n_right, npair, replace=False)
    l2r[pl], r2l[pr] = pr, pl
    dr[pr] = flip_bits(dl[pl], rng.integers(0, 25, npair), rng)      # associated keypoints look alike
    kr["octave"][pr] = kl["octave"][pl]
    desc = np.concatenate([dl, dr])
    occupied = (rng.random(n_left + n_right) < 0.15).astype(np.uint8)
    src_l, src_r = rng.integers(0, n_left, m), rng.integers(0, n_right, m)
    kind = rng.integers(0, 4, m)                                       # 0 left only, 1 right only, 2 both, 3 neither
    inL, inR = np.isin(kind, (0, 2)), np.isin(kind, (1, 2))
    both = kind == 2
    src_r = np.where(both & (l2r[src_l] >= 0), l2r[src_l], src_r)      # "both": the same physical point where possible
    anch = rng.random(m) < 0.8
    px = np.where(anch, kl["x"][src_l] + rng.normal(0, 1.5, m), rng.uniform(0, w, m)).astype(np.float32)
    py = np.where(anch, kl["y"][src_l] + rng.normal(0, 1.5, m), rng.uniform(0, h, m)).astype(np.float32)
    pxr = np.where(anch, kr["x"][src_r] + rng.normal(0, 1.5, m), rng.uniform(0, w, m)).astype(np.float32)
    pyr = np.where(anch, kr["y"][src_r] + rng.normal(0, 1.5, m), rng.uniform(0, h, m)).astype(np.float32)
    lvl = np.clip(kl["octave"][src_l] + rng.integers(0, 2, m), 0, n_levels - 1).astype(np.int32)
    lvr = np.clip(kr["octave"][src_r] + rng.integers(0, 2, m), 0, n_levels - 1).astype(np.int32)
    lvr = np.where(rng.random(m) < 0.05, -1, lvr).astype(np.int32)     # mnTrackScaleLevelR == -1: right search skipped
    d = flip_bits(np.where(inL[:, None], dl[src_l], dr[src_r]), rng.integers(0, 71, m), rng)
    frame = dict(kps_left=kl, kps_right=kr, desc=desc, occupied=occupied, left_to_right=l2r, right_to_left=r2l)
    mp = dict(track_in_view=inL.astype(np.uint8), proj_x=px, proj_y=py, proj_xr=np.zeros(m, np.float32), level=lvl,
              view_cos=rng.uniform(0.5, 1.0, m).astype(np.float32), depth=rng.uniform(0.5, 20.0, m).astype(np.float32),
              has_obs=(rng.random(m) < 0.8).astype(np.uint8), desc=d)
    mpr = dict(track_in_view_r=inR.astype(np.uint8), proj_x_r=pxr, proj_y_r=pyr, level_r=lvr,
               view_cos_r=rng.uniform(0.5, 1.0, m).astype(np.float32))
    return frame, mp, mpr


def projected_points(kps, desc, m, w, h, n_levels, scale_factors, seed=0, bf=47.9, th=7.0, stereo=True):
    """Points of a 'last frame' projected into the current one (input of SearchByProjection(Frame&, const Frame&)):
    most are real keypoints jittered by a small motion, the rest uniform. Returns a dict in the orbx_projected layout."""
    rng = np.random.default_rng(seed + 15485863)
    n = len(kps)
    src = rng.integers(0, n, m)
    anchored = rng.random(m) < 0.8
    u = np.where(anchored, kps["x"][src] + rng.normal(0, 2.0, m), rng.uniform(0, w, m)).astype(np.float32)
    v = np.where(anchored, kps["y"][src] + rng.normal(0, 2.0, m), rng.uniform(0, h, m)).astype(np.float32)
    octave = np.where(anchored, kps["octave"][src], rng.integers(0, n_levels, m)).astype(np.int32)
    mode = rng.integers(0, 3, m)  # forward / backward / neither, src/ORBmatcher.cc:1671-1680
    min_level = np.where(mode == 0, octave, np.where(mode == 1, 0, octave - 1)).astype(np.int32)
    max_level = np.where(mode == 0, -1, np.where(mode == 1, octave, octave + 1)).astype(np.int32)
    radius = (np.float32(th) * np.asarray(scale_factors, np.float32)[octave]).astype(np.float32)
    depth = rng.uniform(0.5, 20.0, m).astype(np.float32)
    d = flip_bits(desc[src], rng.integers(0, 70, m), rng)
    angle = np.where(anchored, kps["angle"][src] + rng.normal(0, 8.0, m), rng.uniform(0, 360, m)) % 360.0
    return dict(u=u, v=v, u_right=(u - np.float32(bf) / depth).astype(np.float32) if stereo else None, radius=radius,
                min_level=min_level, max_level=max_level, angle=angle.astype(np.float32),
                has_obs=(rng.random(m) < 0.9).astype(np.uint8), desc=d)


def feature_vector(n, n_nodes, seed=0):
    """A DBoW2::FeatureVector as CSR: every feature index 0..n-1 falls in exactly one of n_nodes vocabulary nodes
    (ids ascending, sparse), indices inside a node ascending (the order DBoW2 appends them in)."""
    rng = np.random.default_rng(seed + 32452843)
    ids_all = np.sort(rng.choice(10 * n_nodes, n_nodes, replace=False)).astype(np.uint32)
    node_of = rng.integers(0, n_nodes, n)
    order = np.argsort(node_of, kind="stable")
    counts = np.bincount(node_of, minlength=n_nodes)
    keep = counts > 0
    offsets = np.zeros(keep.sum() + 1, np.int32)
    offsets[1:] = np.cumsum(counts[keep])
    return ids_all[keep], offsets, order.astype(np.uint32), node_of


def vocabulary(k=10, depth=4, seed=0, ragged=True):
    """A synthetic DBoW2-style vocabulary tree (k children per node, `depth` levels below the root) in the flat layout of
    orbx_vocabulary: children descriptors are the parent's with a few dozen bits flipped (so descents are meaningful),
    some siblings are exact duplicates (distance ties: the first must win), and with `ragged` a few inner nodes have
    fewer children / become early leaves. Returns a dict in the orbx_vocabulary layout (views.make_vocabulary)."""
    rng = np.random.default_rng(seed + 424243)
    desc = [np.zeros(32, np.uint8)]
    child_lists = [[]]
    level_of = [0]
    frontier = [0]
    for lv in range(1, depth + 1):
        nxt = []
        for node in frontier:
            nc = k
            if ragged and lv > 1 and rng.random() < 0.08:
                nc = int(rng.integers(0, k))          # fewer children, possibly none (an early leaf)
            base = desc[node] if node else rng.integers(0, 256, 32, dtype=np.uint8)
            for c in range(nc):
                d = base.copy() if node else rng.integers(0, 256, 32, dtype=np.uint8)
                for b in rng.choice(256, size=int(rng.integers(8, 48)), replace=False):
                    d[b >> 3] ^= np.uint8(1 << (b & 7))
                if c and rng.random() < 0.1:
                    d = desc[child_lists[node][0]].copy()   # duplicate of the first sibling
                desc.append(d)
                child_lists.append([])
                level_of.append(lv)
                child_lists[node].append(len(desc) - 1)
                nxt.append(len(desc) - 1)
        frontier = nxt
    n = len(desc)
    offsets = np.zeros(n + 1, np.int32)
    children = []
    for i in range(n):
        children.extend(child_lists[i])
        offsets[i + 1] = len(children)
    is_leaf = np.diff(offsets) == 0
    word_id = np.zeros(n, np.uint32)
    word_id[is_leaf] = np.arange(int(is_leaf.sum()), dtype=np.uint32)
    weight = np.where(is_leaf, rng.uniform(0.0, 9.0, n), 0.0)
    weight[is_leaf & (rng.random(n) < 0.05)] = 0.0     # "stopped" words (w > 0 fails, TemplatedVocabulary.h:1154)
    return dict(depth=depth, child_offsets=offsets, children=np.asarray(children, np.uint32),
                descriptors=np.stack(desc), word_id=word_id, weight=weight.astype(np.float64))
