"""BASELINE.json's full-size configurations, checked through size-independent properties plus oracle spot checks
(the oracle finishes a 1280x720 frame in ~0.1 s but a 100k x 100k Hamming sweep would take minutes).

  configs[2]: 1280x720 mono stream, 2000 features, batch of 64 frames       -> batch == frame-by-frame, 3 frames vs oracle
  configs[4]: brute-force 256-bit Hamming 2-NN up to 100k x 100k            -> sampled rows vs numpy, ordering, self-match
"""
import numpy as np
import pytest

from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, synth
from oracle import orbref

pytestmark = pytest.mark.gpu


def test_config3_batch_of_64_frames_1280x720(gpu):
    kinds = ["scene", "noise_blur"]
    distinct = np.stack([synth.make(kinds[s % 2], 720, 1280, s) for s in range(8)])
    imgs = np.ascontiguousarray(np.tile(distinct, (8, 1, 1)))  # 64 frames
    ex = ORBextractor(2000, max_batch=16)                       # 4 pipelined groups of 16
    n, mono, kps, desc = ex.extract_batch(imgs, (0, 1000))       # the monocular lapping of src/Frame.cc:447
    assert n.shape == (64,) and (n >= 1990).all() and (n <= ex.capacity).all()
    # replicas of the same image give the same answer wherever they sit in the batch (no cross-frame state)
    for f in range(8, 64):
        assert n[f] == n[f % 8] and mono[f] == mono[f % 8]
        assert np.array_equal(kps[f, :n[f]], kps[f % 8, :n[f]]) and np.array_equal(desc[f, :n[f]], desc[f % 8, :n[f]])
    # batch == single-frame calls on a fresh handle
    ex1 = ORBextractor(2000)
    for f in (0, 5):
        m1, k1, d1 = ex1(imgs[f], (0, 1000))
        assert m1 == mono[f] and np.array_equal(k1, kps[f, :n[f]]) and np.array_equal(d1, desc[f, :n[f]])
    # and the oracle
    ref = orbref.Extractor(2000)
    for f in (1, 2, 7):
        mr, kr, dr = ref(imgs[f], (0, 1000))
        assert mr == mono[f] and np.array_equal(kr, kps[f, :n[f]]) and np.array_equal(dr, desc[f, :n[f]])
    # lapping split: rows [0, mono) have x outside [0, 1000], rows [mono, n) inside, and the latter are in reverse
    # level order (written from the back, src/ORBextractor.cc:1088-1097)
    f = 0
    x = kps[f, :n[f]]["x"]
    assert (x[:mono[f]] > 1000).all() and (x[mono[f]:] <= 1000).all()
    assert (np.diff(kps[f, mono[f]:n[f]]["octave"]) <= 0).all() and (np.diff(kps[f, :mono[f]]["octave"]) >= 0).all()


def _popcount_rows(q, t):
    qq = q.view(np.uint64)[:, None, :]
    tt = t.view(np.uint64)[None, :, :]
    return np.bitwise_count(qq ^ tt).sum(axis=2).astype(np.int32)


@pytest.mark.parametrize("n", [30000, 100000])
def test_config5_knn2_large_sweep(gpu, n):
    q, t = synth.descriptors(n, 21), synth.descriptors(n, 22)
    t[n // 2] = q[17]                       # a planted exact match
    t[n // 2 + 5] = q[17]                   # and a duplicate of it further down: the lower trainIdx must win
    mt = ORBmatcher()
    idx1, d1, idx2, d2 = mt.knnMatch2(q, t)
    assert idx1.shape == (n,) and (idx1 >= 0).all() and (idx2 >= 0).all() and (idx1 != idx2).all()
    assert (d1 <= d2).all() and (d1 >= 0).all() and (d2 <= 256).all()
    assert idx1[17] == n // 2 and d1[17] == 0 and idx2[17] == n // 2 + 5 and d2[17] == 0
    rng = np.random.default_rng(n)
    rows = np.concatenate([rng.integers(0, n, 96), [0, 17, n - 1]])
    for r0 in range(0, len(rows), 33):
        rr = rows[r0:r0 + 33]
        D = _popcount_rows(q[rr], t)
        order = np.lexsort((np.broadcast_to(np.arange(n), D.shape), D), axis=1)[:, :2]  # by (distance, trainIdx)
        assert np.array_equal(idx1[rr], order[:, 0]) and np.array_equal(idx2[rr], order[:, 1])
        assert np.array_equal(d1[rr], D[np.arange(len(rr)), order[:, 0]])
        assert np.array_equal(d2[rr], D[np.arange(len(rr)), order[:, 1]])
    # self-match: querying the train set against itself returns every row as its own nearest neighbour
    sub = t[:20000]
    i1, e1, _, _ = mt.knnMatch2(sub, sub)
    dup = n // 2 + 5 < 20000
    assert (e1 == 0).all() and (dup or np.array_equal(i1, np.arange(len(sub))))


def test_config4_shapes_at_scale(gpu):
    # 10 000 MapPoints against a 640x480 / 1200-feature frame is covered bit-exactly by
    # tests/test_gpu_matcher.py::test_search_by_projection_map[10000-...]; here: DescriptorDistance at scale
    a, b = synth.descriptors(200000, 1), synth.descriptors(200000, 2)
    d = ORBmatcher().DescriptorDistanceBatch(a, b)
    ref = np.bitwise_count(a.view(np.uint64) ^ b.view(np.uint64)).sum(axis=1)
    assert np.array_equal(d, ref.astype(np.int32))
