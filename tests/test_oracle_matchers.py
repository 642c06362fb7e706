"""Independent cross-checks of the oracle's matcher restatements that have no external pin (DBoW2 / the reference cannot
be built here): each one is compared with a second, differently written (numpy / pure Python) restatement of the same
reference lines, plus the invariants the reference's data structures guarantee. CPU only.

  orbref.search_by_bow / search_by_bow_kf     src/ORBmatcher.cc:230-404, 766-884
  orbref.fuse_match                           src/ORBmatcher.cc:1194-1257 (and the gate-free Sim3 form :1356-1372)
  orbref.search_for_initialization            src/ORBmatcher.cc:618-764
  orbref.bow_transform                        Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1262
  orbref.build_grid                           src/Frame.cc:520-547
"""
import numpy as np

from orb_slam3_fast_b200 import synth, views
from oracle import orbref


def _ham(a, b):
    return int(np.bitwise_count(np.bitwise_xor(a, b)).sum())


def _rot_bin(a1, a2):
    rot = np.float32(a1) - np.float32(a2)
    if rot < 0.0:
        rot = np.float32(rot + np.float32(360.0))
    b = int(np.floor(np.float32(rot * np.float32(1.0 / 30)) + np.float32(0.5)))  # round(): values are >= 0 here
    return 0 if b == 30 else b


def _three_maxima(hist):
    """ComputeThreeMaxima, src/ORBmatcher.cc:1920-1955."""
    max1 = max2 = max3 = 0
    i1 = i2 = i3 = -1
    for i, h in enumerate(hist):
        s = len(h)
        if s > max1:
            max3, max2, max1 = max2, max1, s
            i3, i2, i1 = i2, i1, i
        elif s > max2:
            max3, max2 = max2, s
            i3, i2 = i2, i
        elif s > max3:
            max3, i3 = s, i
    if max2 < 0.1 * np.float32(max1):
        i2 = i3 = -1
    elif max3 < 0.1 * np.float32(max1):
        i3 = -1
    return i1, i2, i3


def _keyframes(seed, n=400, n_nodes=12):
    rng = np.random.default_rng(seed)
    d1 = synth.descriptors(n, seed)
    d2 = synth.flip_bits(d1[rng.permutation(n)], rng.integers(0, 70, n), rng)
    out = []
    for d in (d1, d2):
        kps = np.zeros(n, synth.KP_DTYPE)
        kps["angle"] = rng.uniform(0, 360, n).astype(np.float32)
        kps["octave"] = rng.integers(0, 8, n)
        node_of = d[:, 0].astype(np.int64) % n_nodes * 5 + 2
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        off = np.zeros(len(ids) + 1, np.int32)
        off[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        hm = (rng.random(n) < 0.7).astype(np.uint8)
        out.append(dict(kps=kps, desc=d, hm=hm, ids=ids.astype(np.uint32), off=off, idx=order.astype(np.uint32)))
    return out


def _two_cameras(k, seed, n_nodes=12):
    """A two-camera version of one _keyframes() entry: the rows as they are = the left camera, a noisy copy of every row
    = the right camera (same vocabulary node: byte 0 kept), so that a KeyFrame feature usually has a close left AND a
    close right candidate. Returns (dict, n_left)."""
    rng = np.random.default_rng(seed + 1000)
    n = len(k["kps"])
    dr = k["desc"].copy()
    flips = rng.integers(0, 40, n)
    for i in range(n):
        for bit in rng.choice(np.arange(8, 256), flips[i], replace=False):
            dr[i, bit >> 3] ^= np.uint8(1 << (bit & 7))
    kr = k["kps"].copy()
    kr["angle"] = ((kr["angle"] + rng.normal(0, 25, n)) % 360).astype(np.float32)
    kps = np.concatenate([k["kps"], kr])
    d = np.concatenate([k["desc"], dr])
    node_of = d[:, 0].astype(np.int64) % n_nodes * 5 + 2
    ids, inv = np.unique(node_of, return_inverse=True)
    order = np.argsort(inv, kind="stable")
    off = np.zeros(len(ids) + 1, np.int32)
    off[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
    hm = np.concatenate([k["hm"], (rng.random(n) < 0.7).astype(np.uint8)])
    return dict(kps=kps, desc=d, hm=hm, ids=ids.astype(np.uint32), off=off, idx=order.astype(np.uint32)), n


def _view(k):
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    return orbref.make_keyframe_view(k["kps"], k["desc"], None, k["hm"], k["ids"], k["off"], k["idx"], sf, sf * sf)


def _bow_python(k1, k2, nnratio, check, kf_kf):
    """The two SearchByBoW overloads written once more, straight from the reference text."""
    n_out = len(k1["kps"]) if kf_kf else len(k2["kps"])
    out = [-1] * n_out
    matched2 = [False] * len(k2["kps"])
    hist = [[] for _ in range(30)]
    nm = 0
    nodes2 = {int(i): j for j, i in enumerate(k2["ids"])}
    for a, nid in enumerate(k1["ids"]):
        b = nodes2.get(int(nid))
        if b is None:
            continue
        for i1 in k1["idx"][k1["off"][a]:k1["off"][a + 1]]:
            if not k1["hm"][i1]:
                continue
            best1, best2, bi = 256, 256, -1
            for i2 in k2["idx"][k2["off"][b]:k2["off"][b + 1]]:
                if (kf_kf and (matched2[i2] or not k2["hm"][i2])) or (not kf_kf and out[i2] >= 0):
                    continue
                d = _ham(k1["desc"][i1], k2["desc"][i2])
                if d < best1:
                    best2, best1, bi = best1, d, int(i2)
                elif d < best2:
                    best2 = d
            low = best1 < 50 if kf_kf else best1 <= 50
            if low and np.float32(best1) < np.float32(nnratio) * np.float32(best2):
                if kf_kf:
                    out[i1] = bi
                    matched2[bi] = True
                    key = int(i1)
                else:
                    out[bi] = int(i1)
                    key = bi
                if check:
                    hist[_rot_bin(k1["kps"]["angle"][i1], k2["kps"]["angle"][bi])].append(key)
                nm += 1
    if check:
        keep = _three_maxima(hist)
        for i, h in enumerate(hist):
            if i in keep:
                continue
            for key in h:
                out[key] = -1
                nm -= 1
    return nm, np.asarray(out, np.int32)


def test_search_by_bow_equals_a_python_restatement():
    for seed, nnratio, check in ((0, 0.7, True), (1, 0.9, False), (2, 0.8, True)):
        k1, k2 = _keyframes(seed)
        for kf_kf in (False, True):
            fn = orbref.search_by_bow_kf if kf_kf else orbref.search_by_bow
            n, m = fn(_view(k1), _view(k2), nnratio, check)
            n_p, m_p = _bow_python(k1, k2, nnratio, check, kf_kf)
            assert n == n_p and np.array_equal(m, m_p), (seed, kf_kf)
            assert n_p > 5
            taken = m[m >= 0]
            assert len(np.unique(taken)) == len(taken)      # a feature of the other view is used at most once


def test_bow_transform_descends_to_a_leaf_through_the_recorded_node():
    voc = synth.vocabulary(8, 4, 3)
    off, ch = voc["child_offsets"], voc["children"]
    parent = np.zeros(len(voc["descriptors"]), np.int64)
    for n in range(len(off) - 1):
        parent[ch[off[n]:off[n + 1]]] = n
    leaf_of_word = {int(voc["word_id"][n]): n for n in range(len(off) - 1) if off[n + 1] == off[n]}
    rng = np.random.default_rng(0)
    feats = synth.flip_bits(voc["descriptors"][rng.integers(1, len(parent), 400)], rng.integers(0, 50, 400), rng)
    for levelsup in (0, 2, 4, 9):
        w, wt, nd = orbref.bow_transform(orbref.make_vocabulary(**voc), feats, levelsup)
        for i in range(len(feats)):
            leaf = leaf_of_word[int(w[i])]
            assert wt[i] == voc["weight"][leaf]
            # greedy descent, re-done in numpy: the first child with the least distance at every level
            node, path = 0, [0]
            while off[node + 1] > off[node]:
                kids = ch[off[node]:off[node + 1]]
                d = [_ham(feats[i], voc["descriptors"][k]) for k in kids]
                node = int(kids[int(np.argmin(d))])
                path.append(node)
            assert node == leaf
            lvl = voc["depth"] - levelsup
            assert nd[i] == (path[lvl] if 0 < lvl < len(path) else 0)


def test_fuse_match_equals_a_numpy_restatement():
    rng = np.random.default_rng(4)
    n, m, w, h = 600, 900, 640, 480
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, w, n).astype(np.float32), rng.uniform(0, h, n).astype(np.float32)
    kps["octave"] = rng.integers(0, 8, n)
    desc = synth.descriptors(n, 4)
    ur = np.where(rng.random(n) < 0.5, kps["x"] - rng.uniform(1, 30, n), -1).astype(np.float32)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)
    fr = orbref.make_frame_view(kps, desc, ur, np.zeros(n, np.uint8), g, keep, sf)
    src = rng.integers(0, n, m)
    u = (kps["x"][src] + rng.normal(0, 1.5, m)).astype(np.float32)
    v = (kps["y"][src] + rng.normal(0, 1.5, m)).astype(np.float32)
    lev = np.clip(kps["octave"][src] + rng.integers(-1, 2, m), 0, 7).astype(np.int32)
    radius = (np.float32(3.0) * sf[lev]).astype(np.float32)
    d = synth.flip_bits(desc[src], rng.integers(0, 50, m), rng)
    pur = (u - 4).astype(np.float32)
    pts = orbref.make_projected(u, v, pur, radius, lev - 1, lev, np.zeros(m, np.float32), np.zeros(m, np.uint8), d)
    inv_s2 = (1.0 / (sf * sf)).astype(np.float32)
    for gate in (True, False):
        bi, bd = orbref.fuse_match(fr, inv_s2, pts, gate)
        for i in range(0, m, 7):
            best, bidx = 256, -1
            cx0 = max(0, int(np.floor(np.float32(np.float32(u[i] - radius[i]) * inv_w))))
            cx1 = min(63, int(np.ceil(np.float32(np.float32(u[i] + radius[i]) * inv_w))))
            cy0 = max(0, int(np.floor(np.float32(np.float32(v[i] - radius[i]) * inv_h))))
            cy1 = min(47, int(np.ceil(np.float32(np.float32(v[i] + radius[i]) * inv_h))))
            for cx in range(cx0, cx1 + 1):
                for cy in range(cy0, cy1 + 1):
                    c = cx * 48 + cy
                    for k in items[off[c]:off[c + 1]]:
                        if not (abs(np.float32(kps["x"][k] - u[i])) < radius[i] and abs(np.float32(kps["y"][k] - v[i])) < radius[i]):
                            continue
                        lv = kps["octave"][k]
                        if lv < lev[i] - 1 or lv > lev[i]:
                            continue
                        ex, ey = np.float32(u[i] - kps["x"][k]), np.float32(v[i] - kps["y"][k])
                        if gate:
                            if ur[k] >= 0:
                                er = np.float32(pur[i] - ur[k])
                                e2 = np.float32(np.float32(np.float32(ex * ex) + np.float32(ey * ey)) + np.float32(er * er))
                                if float(np.float32(e2 * inv_s2[lv])) > 7.8:
                                    continue
                            else:
                                e2 = np.float32(np.float32(ex * ex) + np.float32(ey * ey))
                                if float(np.float32(e2 * inv_s2[lv])) > 5.99:
                                    continue
                        dist = _ham(d[i], desc[k])
                        if dist < best:
                            best, bidx = dist, int(k)
            assert bi[i] == bidx and bd[i] == best, (gate, i)
        assert (bi >= 0).sum() > m // 3


def test_search_for_initialization_invariants():
    rng = np.random.default_rng(6)
    n, w, h = 500, 640, 480
    k1 = np.zeros(n, synth.KP_DTYPE)
    k1["x"], k1["y"] = rng.uniform(20, w - 20, n).astype(np.float32), rng.uniform(20, h - 20, n).astype(np.float32)
    k1["octave"] = (rng.random(n) < 0.3).astype(np.int32) * rng.integers(1, 8, n)
    k1["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    d1 = synth.descriptors(n, 6)
    perm = rng.permutation(n)
    k2 = k1[perm].copy()
    k2["x"] += rng.normal(0, 3, n).astype(np.float32)
    k2["y"] += rng.normal(0, 3, n).astype(np.float32)
    d2 = synth.flip_bits(d1[perm], rng.integers(0, 45, n), rng)
    inv_w, inv_h = np.float32(64) / np.float32(w), np.float32(48) / np.float32(h)
    sf = np.float32(1.2) ** np.arange(8, dtype=np.float32)

    def view(k, d):
        off, items = orbref.build_grid(k, 0.0, 0.0, inv_w, inv_h)
        g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
        return orbref.make_frame_view(k, d, None, np.zeros(len(k), np.uint8), g, keep, sf)
    prev = np.stack([k1["x"], k1["y"]], axis=1)
    nm, m12 = orbref.search_for_initialization(view(k1, d1), view(k2, d2), prev, 30, 0.9, True)
    matched = np.nonzero(m12 >= 0)[0]
    assert nm == len(matched) and nm > 50
    assert (k1["octave"][matched] == 0).all() and (k2["octave"][m12[matched]] == 0).all()   # level 0 on both sides
    assert len(np.unique(m12[matched])) == len(matched)                                      # vnMatches21 is a map
    for i in matched:                                                                         # inside the window, TH_LOW
        j = m12[i]
        assert abs(k2["x"][j] - prev[i, 0]) < 30 and abs(k2["y"][j] - prev[i, 1]) < 30
        assert _ham(d1[i], d2[j]) <= 50
    # without the rotation check nothing is removed afterwards: at least as many matches
    nm2, _ = orbref.search_for_initialization(view(k1, d1), view(k2, d2), prev, 30, 0.9, False)
    assert nm2 >= nm


def test_build_grid_is_the_host_mirror():
    rng = np.random.default_rng(8)
    kps = np.zeros(700, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(-20, 660, 700).astype(np.float32), rng.uniform(-20, 500, 700).astype(np.float32)
    inv_w, inv_h = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    off_h, items_h = views.assign_features_to_grid(kps, 0.0, 0.0, inv_w, inv_h)
    assert np.array_equal(off, off_h) and np.array_equal(items, items_h[:off_h[-1]])
    for c in range(64 * 48):
        assert (np.diff(items[off[c]:off[c + 1]]) > 0).all()        # push_back order = ascending index


def _features_in_area(kps, off, items, x, y, r, min_level, max_level, inv_w, inv_h):
    """Frame::GetFeaturesInArea (src/Frame.cc:765-831), min corner (0, 0), written from the reference text."""
    f32 = np.float32
    x0 = max(0, int(np.floor(f32(f32(x - r) * inv_w))))
    if x0 >= 64:
        return []
    x1 = min(63, int(np.ceil(f32(f32(x + r) * inv_w))))
    if x1 < 0:
        return []
    y0 = max(0, int(np.floor(f32(f32(y - r) * inv_h))))
    if y0 >= 48:
        return []
    y1 = min(47, int(np.ceil(f32(f32(y + r) * inv_h))))
    if y1 < 0:
        return []
    check = min_level > 0 or max_level >= 0
    out = []
    for cx in range(x0, x1 + 1):
        for cy in range(y0, y1 + 1):
            c = cx * 48 + cy
            for k in items[off[c]:off[c + 1]]:
                if check:
                    if kps["octave"][k] < min_level:
                        continue
                    if max_level >= 0 and kps["octave"][k] > max_level:
                        continue
                if abs(f32(kps["x"][k] - x)) < r and abs(f32(kps["y"][k] - y)) < r:
                    out.append(int(k))
    return out


def test_search_by_projection_map_equals_a_python_restatement():
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFarPoints), serial MapPoint
    order (src/ORBmatcher.cc:42-146): the row SURVEY.md §8 calls a13."""
    f32 = np.float32
    rng = np.random.default_rng(3)
    n, w, h = 500, 640, 480
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
    kps["octave"] = rng.integers(0, 8, n)
    desc = synth.descriptors(n, 3)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    for stereo, th, nnratio in ((True, 1.0, 0.8), (False, 3.0, 0.8), (True, 5.0, 0.6)):
        ur = np.where(rng.random(n) < 0.7, kps["x"] - rng.uniform(1, 40, n), -1).astype(f32) if stereo else None
        occ = (rng.random(n) < 0.1).astype(np.uint8)
        fr = orbref.make_frame_view(kps, desc, ur, occ, g, keep, sf)
        mp = synth.local_map(kps, desc, 1500, w, h, 8, int(th * 10))
        n_o, a_o = orbref.search_by_projection_map(fr, orbref.make_mappoints(**mp), th, nnratio, True, 15.0)
        # ---- the same loop in Python ----
        assign = [-1] * n
        blocked = occ.astype(bool).copy()        # mvpMapPoints[idx] with Observations() > 0
        nm = 0
        for i in range(len(mp["proj_x"])):
            if not mp["track_in_view"][i] or mp["depth"][i] > f32(15.0):
                continue
            L = int(mp["level"][i])
            r = f32(2.5) if float(mp["view_cos"][i]) > 0.998 else f32(4.0)
            if th != 1.0:
                r = f32(r * f32(th))
            rs = f32(r * sf[L])
            best, best2, bl, bl2, bi = 256, 256, -1, -1, -1
            for k in _features_in_area(kps, off, items, mp["proj_x"][i], mp["proj_y"][i], rs, L - 1, L, inv_w, inv_h):
                if blocked[k]:
                    continue
                if stereo and ur[k] > 0 and abs(f32(mp["proj_xr"][i] - ur[k])) > rs:
                    continue
                d = _ham(mp["desc"][i], desc[k])
                if d < best:
                    best2, best, bl2, bl, bi = best, d, bl, int(kps["octave"][k]), k
                elif d < best2:
                    bl2, best2 = int(kps["octave"][k]), d
            if best <= 100:
                if bl == bl2 and f32(best) > f32(f32(nnratio) * f32(best2)):
                    continue
                assign[bi] = i
                blocked[bi] = bool(mp["has_obs"][i])
                nm += 1
        assert nm == n_o and np.array_equal(np.asarray(assign, np.int32), a_o), (stereo, th)
        assert n_o > 30


def _stereo_python(ex_l, ex_r, kl, dl, kr, dr, mbf, mb):
    """Frame::ComputeStereoMatches, src/Frame.cc:921-1084, written over numpy slices of the oracle's pyramid levels."""
    import math
    f32 = np.float32
    n = len(kl)
    u_right = np.full(n, -1.0, f32)
    depth = np.full(n, -1.0, f32)
    n_rows = ex_l.level_dims(0)[1]
    sf, inv_sf = ex_l.scale, ex_l.inv_scale
    pyr_l = [ex_l.level_image(l).astype(np.int32) for l in range(ex_l.nlevels)]
    pyr_r = [ex_r.level_image(l).astype(np.int32) for l in range(ex_r.nlevels)]
    # row table: right keypoint iR is a candidate of row y  <=>  floor(y_R - r) <= y <= ceil(y_R + r), in index order
    r = f32(2.0) * sf[kr["octave"]]
    lo = np.floor((kr["y"] - r).astype(f32)).astype(np.int64)
    hi = np.ceil((kr["y"] + r).astype(f32)).astype(np.int64)
    live = ~((kr["x"] == 0) & (kr["y"] == 0))
    max_d = f32(f32(mbf) / f32(mb))
    dist_idx = []
    for i in range(n):
        u_l, v_l, lev = f32(kl["x"][i]), f32(kl["y"][i]), int(kl["octave"][i])
        row = int(v_l)
        if not 0 <= row < n_rows:
            continue
        cand = np.flatnonzero(live & (lo <= row) & (row <= hi))
        if len(cand) == 0:
            continue
        min_u, max_u = f32(u_l - max_d), u_l
        if max_u < 0:
            continue
        ok = cand[(np.abs(kr["octave"][cand].astype(np.int64) - lev) <= 1) & (kr["x"][cand] >= min_u)
                  & (kr["x"][cand] <= max_u)]
        if len(ok) == 0:
            continue
        d = np.bitwise_count(np.bitwise_xor(dr[ok], dl[i][None, :])).sum(1).astype(np.int64)
        j = int(np.argmin(d))  # first minimum == the serial strict-< scan
        if not (d[j] < 100 and d[j] < 75):
            continue
        rnd = lambda x: math.floor(float(x) + 0.5)  # std::round for the non-negative values here
        s = inv_sf[lev]
        su_l, sv_l, su_r0 = rnd(f32(u_l * s)), rnd(f32(v_l * s)), rnd(f32(f32(kr["x"][ok[j]]) * s))
        w = L = 5
        img_l, img_r = pyr_l[lev], pyr_r[lev]
        if su_r0 + L - w < 0 or su_r0 + L + w + 1 >= img_r.shape[1]:
            continue
        patch = img_l[sv_l - w:sv_l + w + 1, su_l - w:su_l + w + 1]
        sad = np.array([np.abs(patch - img_r[sv_l - w:sv_l + w + 1, su_r0 + inc - w:su_r0 + inc + w + 1]).sum()
                        for inc in range(-L, L + 1)], np.int64)
        b = int(np.argmin(sad))
        if b == 0 or b == 2 * L:
            continue
        d1, d2, d3 = f32(sad[b - 1]), f32(sad[b]), f32(sad[b + 1])
        with np.errstate(all="ignore"):
            delta = f32(f32(d1 - d3) / f32(f32(2.0) * f32(f32(d1 + d3) - f32(f32(2.0) * d2))))
            if delta < -1 or delta > 1:
                continue
            best_ur = f32(sf[lev] * f32(f32(f32(su_r0) + f32(b - L)) + delta))
            disparity = f32(u_l - best_ur)
        if disparity >= 0 and disparity < max_d:
            if disparity <= 0:
                disparity = f32(0.01)
                best_ur = f32(float(u_l) - 0.01)
            depth[i] = f32(f32(mbf) / disparity)
            u_right[i] = best_ur
            dist_idx.append((int(sad[b]), i))
    if not dist_idx:
        return 0, u_right, depth
    dist_idx.sort()
    th = f32(f32(f32(1.5) * f32(1.4)) * f32(dist_idx[len(dist_idx) // 2][0]))
    kept = len(dist_idx)
    for sad_b, i in reversed(dist_idx):
        if f32(sad_b) < th:
            break
        u_right[i] = depth[i] = -1
        kept -= 1
    return kept, u_right, depth


def test_stereo_match_equals_a_python_restatement():
    for (w, h, nf, seed, kind) in ((320, 240, 400, 3, "scene"), (400, 300, 600, 8, "scene")):
        left, right, _ = synth.stereo_pair(h, w, seed, kind=kind)
        ex_l, ex_r = orbref.Extractor(nf), orbref.Extractor(nf)
        _, kl, dl = ex_l(left)
        _, kr, dr = ex_r(right)
        mbf, mb = 0.8 * w * 0.11, 0.11
        n_o, ur_o, dp_o = orbref.stereo_match(ex_l, ex_r, kl, dl, kr, dr, float(mbf), float(mb))
        n_p, ur_p, dp_p = _stereo_python(ex_l, ex_r, kl, dl, kr, dr, mbf, mb)
        assert n_o > 20, "case must exercise the match path"
        assert n_p == n_o
        assert np.array_equal(ur_p.view(np.uint32), ur_o.view(np.uint32))
        assert np.array_equal(dp_p.view(np.uint32), dp_o.view(np.uint32))


def test_search_by_projection_frame_equals_a_python_restatement():
    """The matching stage of ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
    (src/ORBmatcher.cc:1574-1722) over pre-projected points: window by forward / backward / neither level range, the
    already-matched rule, the uRight gate, strict-< best, rotation-histogram pruning."""
    f32 = np.float32
    rng = np.random.default_rng(17)
    n, w, h = 600, 640, 480
    kps = np.zeros(n, synth.KP_DTYPE)
    kps["x"], kps["y"] = rng.uniform(0, w, n).astype(f32), rng.uniform(0, h, n).astype(f32)
    kps["octave"] = rng.integers(0, 8, n)
    kps["angle"] = rng.uniform(0, 360, n).astype(f32)
    desc = synth.descriptors(n, 17)
    sf = f32(1.2) ** np.arange(8, dtype=f32)
    inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
    off, items = orbref.build_grid(kps, 0.0, 0.0, inv_w, inv_h)
    g, keep = orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h)
    for stereo, th, check, max_dist in ((True, 7.0, True, 100), (False, 15.0, True, 100), (True, 7.0, False, 60)):
        ur = np.where(rng.random(n) < 0.7, kps["x"] - rng.uniform(1, 40, n), -1).astype(f32) if stereo else None
        occ = (rng.random(n) < 0.1).astype(np.uint8)
        fr = orbref.make_frame_view(kps, desc, ur, occ, g, keep, sf)
        pts = synth.projected_points(kps, desc, 900, w, h, 8, sf, seed=int(th), th=th, stereo=stereo)
        n_o, a_o = orbref.search_by_projection_frame(fr, orbref.make_projected(**pts), max_dist, check)
        assign = np.full(n, -1, np.int32)
        blocked = occ.astype(bool).copy()
        hist = [[] for _ in range(30)]
        accepted = 0
        for i in range(len(pts["u"])):
            best, bi = 256, -1
            for k in _features_in_area(kps, off, items, pts["u"][i], pts["v"][i], pts["radius"][i],
                                       int(pts["min_level"][i]), int(pts["max_level"][i]), inv_w, inv_h):
                if blocked[k]:
                    continue
                if stereo and ur[k] > 0 and abs(f32(pts["u_right"][i] - ur[k])) > pts["radius"][i]:
                    continue
                d = _ham(pts["desc"][i], desc[k])
                if d < best:
                    best, bi = d, k
            if best <= max_dist:
                assign[bi] = i
                blocked[bi] = bool(pts["has_obs"][i])
                accepted += 1
                if check:
                    hist[_rot_bin(pts["angle"][i], kps["angle"][bi])].append(bi)
        keep_bins = ()
        if check:
            keep_bins = _three_maxima(hist)
            for b in range(30):
                if b not in keep_bins:
                    for k in hist[b]:
                        assign[k] = -1
        assert np.array_equal(assign, a_o), (stereo, th, check)
        # the reference's counter: +1 per accepted point, -1 per pruned histogram entry (an entry whose keypoint was
        # re-assigned later is still counted once per entry)
        nm = accepted - (sum(len(hist[b]) for b in range(30) if b not in keep_bins) if check else 0)
        assert nm == n_o and n_o > 30


def test_search_for_triangulation_equals_a_python_restatement():
    """ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:886-1103) with Pinhole::epipolarConstrain
    (src/CameraModels/Pinhole.cpp:122-149) given F12. Written per shared vocabulary node as: the winner for idx1 is the
    LAST candidate, in node order, that attains the minimum distance among the candidates passing every gate with
    distance <= TH_LOW -- the closed form of the reference's `dist > bestDist -> continue` running scan."""
    f32 = np.float32
    w, h = 640, 400
    left, right, _ = synth.stereo_pair(h, w, 5, d_min=5, d_max=30)
    e1, e2 = orbref.Extractor(1200), orbref.Extractor(1200)
    _, k1, d1 = e1(left)
    _, k2, d2 = e2(right)
    rng = np.random.default_rng(5)

    def featvec(desc):
        node_of = (desc[:, 0].astype(np.int64) >> 4) * 16 + (desc[:, 1].astype(np.int64) >> 4)
        ids, inv = np.unique(node_of, return_inverse=True)
        order = np.argsort(inv, kind="stable")
        offsets = np.zeros(len(ids) + 1, np.int32)
        offsets[1:] = np.cumsum(np.bincount(inv, minlength=len(ids)))
        return ids.astype(np.uint32), offsets, order.astype(np.uint32)

    sf, s2 = e1.scale, e1.sigma2
    kfs = []
    for k, d in ((k1, d1), (k2, d2)):
        ids, off, idx = featvec(d)
        ur = np.where(rng.random(len(k)) < 0.5, k["x"] - 10, -1).astype(f32)
        hm = (rng.random(len(k)) < 0.2).astype(np.uint8)
        kfs.append((k, d, ur, hm, ids, off, idx, sf, s2))
    # near-horizontal translation with a slight tilt so that every term of the line equation is exercised
    F12 = np.array([[1e-7, 2e-6, -3e-4], [-2e-6, 1e-7, -1], [4e-4, 1, 2e-2]], f32)
    for only_stereo, coarse, check, ep in ((False, False, True, (1e6, 200.0)), (True, False, True, (1e6, 200.0)),
                                           (False, True, False, (320.0, 200.0)), (False, False, False, (100.0, 150.0))):
        n_o, m_o = orbref.search_for_triangulation(orbref.make_keyframe_view(*kfs[0]),
                                                   orbref.make_keyframe_view(*kfs[1]), F12, ep, only_stereo, coarse,
                                                   check)
        (ka, da, ua, ha, ia, oa, xa, _, _), (kb, db, ub, hb, ib, ob, xb, _, _) = kfs
        m12 = np.full(len(ka), -1, np.int32)
        hist = [[] for _ in range(30)]
        nm = 0
        Ff = F12.ravel()
        for node in np.intersect1d(ia, ib):
            a, b = int(np.searchsorted(ia, node)), int(np.searchsorted(ib, node))
            c2 = xb[ob[b]:ob[b + 1]].astype(np.int64)
            c2 = c2[hb[c2] == 0]
            st2 = ub[c2] >= 0
            if only_stereo:
                c2, st2 = c2[st2], st2[st2]
            for i1 in xa[oa[a]:oa[a + 1]].astype(np.int64):
                st1 = ua[i1] >= 0
                if ha[i1] or (only_stereo and not st1) or len(c2) == 0:
                    continue
                dist = np.bitwise_count(np.bitwise_xor(db[c2], da[i1][None, :])).sum(1).astype(np.int64)
                x2, y2 = kb["x"][c2], kb["y"][c2]
                ok = dist <= 50
                ex, ey = (f32(ep[0]) - x2).astype(f32), (f32(ep[1]) - y2).astype(f32)
                near_epipole = (ex * ex + ey * ey).astype(f32) < (f32(100) * sf[kb["octave"][c2]]).astype(f32)
                ok &= ~(near_epipole & ~st2 & (not st1))
                if not coarse:
                    x1, y1 = f32(ka["x"][i1]), f32(ka["y"][i1])
                    la = f32(f32(f32(x1 * Ff[0]) + f32(y1 * Ff[3])) + Ff[6])
                    lb = f32(f32(f32(x1 * Ff[1]) + f32(y1 * Ff[4])) + Ff[7])
                    lc = f32(f32(f32(x1 * Ff[2]) + f32(y1 * Ff[5])) + Ff[8])
                    num = ((la * x2).astype(f32) + (lb * y2).astype(f32)).astype(f32) + lc
                    den = f32(f32(la * la) + f32(lb * lb))
                    if den == 0:
                        continue
                    dsqr = ((num * num).astype(f32) / den).astype(f32)
                    ok &= dsqr.astype(np.float64) < 3.84 * s2[kb["octave"][c2]].astype(np.float64)
                if not ok.any():
                    continue
                best = dist[ok].min()
                j = int(np.flatnonzero(ok & (dist == best))[-1])
                m12[i1] = c2[j]
                nm += 1
                if check:
                    hist[_rot_bin(ka["angle"][i1], kb["angle"][c2[j]])].append(int(i1))
        if check:
            keep_bins = _three_maxima(hist)
            for bn in range(30):
                if bn not in keep_bins:
                    for i1 in hist[bn]:
                        m12[i1] = -1
                        nm -= 1
        assert n_o > 20, (only_stereo, coarse, n_o)
        assert nm == n_o and np.array_equal(m12, m_o), (only_stereo, coarse, check)


def test_search_for_initialization_equals_a_python_restatement():
    """ORBmatcher::SearchForInitialization, serial order (src/ORBmatcher.cc:618-764): level-0 keypoints only, window via
    GetFeaturesInArea(level 0..0), candidates already matched at a distance <= dist are skipped, ratio test in float,
    a re-matched F2 keypoint un-matches its previous F1 partner, rotation-histogram pruning counts only live entries."""
    f32 = np.float32
    for seed, window, nnratio, check in ((6, 30, 0.9, True), (7, 60, 0.9, False), (8, 100, 0.7, True)):
        rng = np.random.default_rng(seed)
        n, w, h = 500, 640, 480
        k1 = np.zeros(n, synth.KP_DTYPE)
        k1["x"], k1["y"] = rng.uniform(20, w - 20, n).astype(f32), rng.uniform(20, h - 20, n).astype(f32)
        k1["octave"] = (rng.random(n) < 0.3).astype(np.int32) * rng.integers(1, 8, n)
        k1["angle"] = rng.uniform(0, 360, n).astype(f32)
        # few prototypes, individually perturbed: several F1 points compete for the same F2 point at different distances
        d1 = synth.flip_bits(synth.descriptors(n, seed, 40), rng.integers(0, 12, n), rng)
        perm = rng.permutation(n)
        k2 = k1[perm].copy()
        k2["x"] += rng.normal(0, 3, n).astype(f32)
        k2["y"] += rng.normal(0, 3, n).astype(f32)
        k2["angle"] = ((k2["angle"] + rng.normal(0, 20, n)) % 360).astype(f32)
        d2 = synth.flip_bits(d1[perm], rng.integers(0, 30, n), rng)
        inv_w, inv_h = f32(64) / f32(w), f32(48) / f32(h)
        sf = f32(1.2) ** np.arange(8, dtype=f32)
        grids = []
        for k in (k1, k2):
            off, items = orbref.build_grid(k, 0.0, 0.0, inv_w, inv_h)
            grids.append((off, items) + orbref.make_grid(off, items, 0.0, 0.0, inv_w, inv_h))
        v1 = orbref.make_frame_view(k1, d1, None, np.zeros(n, np.uint8), grids[0][2], grids[0][3], sf)
        v2 = orbref.make_frame_view(k2, d2, None, np.zeros(n, np.uint8), grids[1][2], grids[1][3], sf)
        prev = np.stack([k1["x"], k1["y"]], axis=1)
        n_o, m_o = orbref.search_for_initialization(v1, v2, prev, window, nnratio, check)
        # ---- the same loop in Python ----
        big = 2 ** 31 - 1
        m12, m21, matched_dist = [-1] * n, [-1] * n, [big] * n
        hist = [[] for _ in range(30)]
        nm = rematched = 0
        for i1 in range(n):
            if k1["octave"][i1] > 0:
                continue
            cands = _features_in_area(k2, grids[1][0], grids[1][1], prev[i1, 0], prev[i1, 1], f32(window), 0, 0,
                                      inv_w, inv_h)
            best, best2, bi = big, big, -1
            for i2 in cands:
                d = _ham(d1[i1], d2[i2])
                if matched_dist[i2] <= d:
                    continue
                if d < best:
                    best2, best, bi = best, d, i2
                elif d < best2:
                    best2 = d
            if best <= 50 and f32(best) < f32(f32(best2) * f32(nnratio)):
                if m21[bi] >= 0:
                    m12[m21[bi]] = -1
                    nm -= 1
                    rematched += 1
                m12[i1], m21[bi], matched_dist[bi] = bi, i1, best
                nm += 1
                if check:
                    hist[_rot_bin(k1["angle"][i1], k2["angle"][bi])].append(i1)
        if check:
            keep_bins = _three_maxima(hist)
            for b in range(30):
                if b not in keep_bins:
                    for i1 in hist[b]:
                        if m12[i1] >= 0:
                            m12[i1] = -1
                            nm -= 1
        assert nm == n_o and np.array_equal(np.asarray(m12, np.int32), m_o), (seed, window, check)
        assert n_o > 30 and (rematched > 0 or window < 60), "the re-match path must be exercised"
