"""Frame::ComputeBoW (src/Frame.cc:846-851) and KeyFrame::ComputeBoW (src/KeyFrame.cc:98-107) — SURVEY.md §8(f) rank 2.
Three parties over DBoW2's OWN TemplatedVocabulary<FORB> (compiled where it lies, fed synthetic trees through its own
loadFromTextFile):
  * the reference's own text of the two functions on stand-in Frame / KeyFrame objects (oracle/ref_stubs/bow_world.h;
    oracle/_ref/liborbref_dbow2_src.so),
  * the oracle's per-feature descent + the accumulation DBoW2 prescribes, put together here,
  * the drop-in bodies of shim/FrameBoW_orbx.cc linked in place of the reference's two functions, with the orbm C ABI
    answered by the oracle (oracle/_ref/libshim_bow_world.so, CPU) and by liborbx.so (…_gpu.so, -m gpu).
mBowVec (words and the bits of every double) and mFeatVec must be identical, including the "already computed" guards."""
import os

import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref, refsrc

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
_CPU_WORLD = os.path.join(_REF, "libshim_bow_world.so")
_GPU_WORLD = os.path.join(_REF, "libshim_bow_world_gpu.so")
pytestmark = pytest.mark.skipif(not refsrc.vocabulary_available(), reason="oracle/_ref not built")

# (k, depth, ragged, seed): ComputeBoW files features under the node 4 levels above the leaves; depth 4 and 3 put that at
# (or above) the root, which DBoW2 defines as node 0 (TemplatedVocabulary.h:1227). Ragged trees only where that holds: a
# leaf above the filing level would leave DBoW2's nid unassigned (:1149, undefined there).
CASES = [(10, 4, True, 0), (10, 3, False, 1), (6, 5, False, 2), (4, 6, False, 3), (10, 4, False, 4)]


def _case(tmp_path, k, depth, ragged, seed, n_feats=900):
    voc = synth.vocabulary(k, depth, seed, ragged)
    path = str(tmp_path / "voc.txt")
    refsrc.write_vocabulary_text(voc, k, path)
    rng = np.random.default_rng(seed + 77)
    leaves = np.flatnonzero(np.diff(voc["child_offsets"]) == 0)
    n_leaf = n_feats * 3 // 4
    feats = np.concatenate([synth.flip_bits(voc["descriptors"][rng.choice(leaves, n_leaf)], rng.integers(0, 40, n_leaf), rng),
                            synth.descriptors(n_feats - n_leaf, seed + 50)])
    return voc, path, feats


def _expected(voc, feats):
    word, weight, node = orbref.bow_transform(orbref.make_vocabulary(**voc), feats, 4)
    filed = weight > 0
    acc = {}
    for w, x in zip(word[filed], weight[filed]):              # BowVector::addWeight in feature order (:1156)
        acc[int(w)] = acc.get(int(w), 0.0) + float(x)
    norm = 0.0
    for w in sorted(acc):                                      # BowVector::normalize(L1) (:1198)
        norm += abs(acc[w])
    words = np.array(sorted(acc), np.uint32)
    values = np.array([acc[w] / norm for w in sorted(acc)]) if norm > 0 else np.array([acc[w] for w in sorted(acc)])
    return np.where(filed, node, 0xffffffff).astype(np.uint32), words, values


def _check(rv, voc, feats):
    node_w, words_w, values_w = _expected(voc, feats)
    assert (node_w != 0xffffffff).sum() > len(feats) // 2
    for which in (0, 1, 2):          # Frame, KeyFrame, KeyFrame with a stale mBowVec and no mFeatVec: all computed afresh
        node, words, values = rv.compute_bow(feats, which)
        assert np.array_equal(node, node_w), which
        assert np.array_equal(words, words_w) and values.tobytes() == values_w.tobytes(), which
    node, words, values = rv.compute_bow(feats, 3)   # Frame whose mBowVec is not empty: left alone (src/Frame.cc:847)
    assert (node == 0xffffffff).all() and words.tolist() == [7] and values.tolist() == [0.25]
    node, words, values = rv.compute_bow(feats[:0], 0)
    assert len(node) == 0 and len(words) == 0


@pytest.mark.parametrize("k,depth,ragged,seed", CASES)
def test_reference_compute_bow_equals_the_oracle(tmp_path, k, depth, ragged, seed):
    voc, path, feats = _case(tmp_path, k, depth, ragged, seed)
    _check(refsrc.ReferenceVocabulary(path), voc, feats)


@pytest.mark.skipif(not os.path.exists(_CPU_WORLD), reason="oracle/_ref/libshim_bow_world.so not built")
@pytest.mark.parametrize("k,depth,ragged,seed", CASES)
def test_shim_compute_bow(tmp_path, k, depth, ragged, seed):
    voc, path, feats = _case(tmp_path, k, depth, ragged, seed)
    _check(refsrc.ReferenceVocabulary(path, world=_CPU_WORLD), voc, feats)


@pytest.mark.skipif(not os.path.exists(_CPU_WORLD), reason="oracle/_ref/libshim_bow_world.so not built")
def test_shim_compute_bow_switches_vocabularies(tmp_path):
    """one thread, two vocabularies alternating: the per-thread upload must follow the object it is asked about"""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    va, pa, fa = _case(tmp_path / "a", 10, 4, False, 5, 300)
    vb, pb, fb = _case(tmp_path / "b", 6, 5, False, 6, 300)
    ra = refsrc.ReferenceVocabulary(pa, world=_CPU_WORLD)
    rb = refsrc.ReferenceVocabulary(pb, world=_CPU_WORLD)
    for _ in range(2):
        for rv, voc, feats in ((ra, va, fa), (rb, vb, fb)):
            node_w, words_w, values_w = _expected(voc, feats)
            node, words, values = rv.compute_bow(feats, 0)
            assert np.array_equal(node, node_w) and np.array_equal(words, words_w) and values.tobytes() == values_w.tobytes()


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(_GPU_WORLD), reason="oracle/_ref/libshim_bow_world_gpu.so not built")
@pytest.mark.parametrize("k,depth,ragged,seed", CASES)
def test_gpu_shim_compute_bow(gpu, tmp_path, k, depth, ragged, seed):
    """the drop-in bodies on the CUDA library (orbm_set_vocabulary + orbm_bow_transform of liborbx.so)"""
    import subprocess
    out = subprocess.run(["ldd", _GPU_WORLD], capture_output=True, text=True).stdout
    assert "liborbx.so" in out and "not found" not in out
    voc, path, feats = _case(tmp_path, k, depth, ragged, seed)
    _check(refsrc.ReferenceVocabulary(path, world=_GPU_WORLD), voc, feats)


@pytest.mark.skipif(not os.path.exists(_CPU_WORLD), reason="oracle/_ref/libshim_bow_world.so not built")
def test_shim_compute_bow_from_three_threads(tmp_path):
    """Frame::ComputeBoW runs on the Tracking thread, KeyFrame::ComputeBoW on LocalMapping's and LoopClosing's: three
    threads at once (ctypes releases the GIL during the call), each alternating between two vocabularies — the matcher
    context and the record of which vocabulary it holds are per thread, so nobody may see another thread's tree."""
    import threading
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    cases = []
    for sub, (k, depth, seed) in (("a", (10, 4, 11)), ("b", (6, 5, 12))):
        voc, path, feats = _case(tmp_path / sub, k, depth, False, seed, 400)
        cases.append((refsrc.ReferenceVocabulary(path, world=_CPU_WORLD), _expected(voc, feats), feats))
    errors = []

    def worker(t):
        try:
            for it in range(6):
                rv, (node_w, words_w, values_w), feats = cases[(t + it) % 2]
                node, words, values = rv.compute_bow(feats, (t + it) % 3 if (t + it) % 3 < 2 else 1)
                if not (np.array_equal(node, node_w) and np.array_equal(words, words_w) and values.tobytes() == values_w.tobytes()):
                    errors.append((t, it))
        except Exception as e:  # noqa: BLE001 - report whatever a thread died of
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(3)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
