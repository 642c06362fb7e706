"""The oracle against the reference's OWN extractor source. oracle/_ref/liborbref_src.so is /root/reference/src/
ORBextractor.cc compiled where it lies (oracle/Makefile, target `ref`) against stand-in OpenCV / TBB headers
(oracle/ref_stubs): types with OpenCV's semantics, TBB executed serially, and the five image primitives
(resize, GaussianBlur, FAST, copyMakeBorder, fastAtan2) forwarded to the oracle's restatements that
test_oracle_primitives.py pins byte for byte to the real OpenCV 4.13 kernels. Everything else that runs here is the
reference's code: the constructor tables, ComputePyramid, the cell loop with the iniTh -> minTh retry,
DistributeOctTree / DivideNode / the sorted expansion, IC_Angle, computeOrbDescriptor, the mono / stereo assembly of
operator(). CPU only; skipped where the reference tree was not available at build time."""
import numpy as np
import pytest

from orb_slam3_fast_b200 import synth
from oracle import orbref, refsrc

pytestmark = pytest.mark.skipif(not refsrc.available(), reason="oracle/_ref not built (no /root/reference at build time)")


def _same(params, img, lapping):
    r = refsrc.ReferenceExtractor(*params)
    o = orbref.Extractor(*params)
    mono_r, k_r, d_r = r(img, lapping)
    mono_o, k_o, d_o = o(img, lapping)
    assert mono_r == mono_o and len(k_r) == len(k_o)
    assert np.array_equal(k_r, k_o), np.nonzero(k_r != k_o)[0][:5]
    assert np.array_equal(d_r, d_o)
    return len(k_r)


@pytest.mark.parametrize("params", [(1000, 1.2, 8, 20, 7), (1200, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7),
                                    (500, 1.5, 5, 15, 5), (300, 2.0, 3, 20, 7), (1500, 1.2, 10, 20, 7)])
def test_constructor_tables(params):
    """src/ORBextractor.cc:408-469: scale / sigma tables, per-level quotas, umax."""
    r, o = refsrc.ReferenceExtractor(*params), orbref.Extractor(*params)
    for a, b in zip(r.tables(), (o.scale, o.inv_scale, o.sigma2, o.inv_sigma2, o.features_per_level, o.umax)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("kind,w,h,nfeat,lapping,seed", [
    ("scene", 640, 480, 1000, (0, 0), 0), ("scene", 640, 480, 1000, (0, 1000), 1),   # BASELINE configs[0], both lappings
    ("scene", 752, 480, 1200, (0, 0), 2), ("noise_blur", 752, 480, 1200, (0, 0), 3),  # configs[1]
    ("scene", 1280, 720, 2000, (0, 1000), 4),                                         # configs[2]: x > 1000 goes to the front
    ("uniform_noise", 640, 480, 1000, (0, 0), 5),                                     # 65 k candidates into the quadtree
    ("low_contrast40", 640, 480, 1000, (0, 0), 6),                                    # minThFAST retry in most cells
    ("scene", 241, 241, 300, (0, 0), 7), ("scene", 333, 517, 700, (100, 200), 8)])    # minimum size; partial lapping
def test_extractor_equals_the_reference_source(kind, w, h, nfeat, lapping, seed):
    img = synth.low_contrast(h, w, seed, 40) if kind == "low_contrast40" else synth.make(kind, h, w, seed)
    n = _same((nfeat, 1.2, 8, 20, 7), img, lapping)
    assert n > 0


def test_stereo_pair_and_other_parameters():
    left, right, _ = synth.stereo_pair(480, 752, 11)
    for img in (left, right):
        _same((1200, 1.2, 8, 20, 7), img, (0, 0))
    img = synth.scene(480, 640, 12)
    _same((800, 1.5, 5, 15, 5), img, (0, 0))
    _same((600, 2.0, 3, 20, 7), img, (0, 0))
    _same((1000, 1.2, 8, 20, 20), img, (0, 0))   # iniTh == minTh: the retry is the same call again


def test_degenerate_images():
    assert _same((1000, 1.2, 8, 20, 7), synth.constant(480, 640), (0, 0)) == 0       # N = 0: released descriptors
    assert _same((1000, 1.2, 8, 20, 7), synth.low_contrast(480, 640, 1, 12), (0, 0)) == 0  # no corner at either threshold
    r = refsrc.ReferenceExtractor(1000)
    mono, k, d = r(None)
    assert mono == -1 and len(k) == 0                                                # :1021, empty image
    assert orbref.Extractor(1000)(None)[0] == -1


def test_quadtree_of_the_reference_source_on_the_oracles_candidates():
    """DistributeOctTree (:557-757) called directly: fed with the candidate list the oracle built for each level, the
    reference's quadtree must return the oracle's kept keypoints in the oracle's order."""
    img = synth.uniform_noise(480, 640, 21)
    o = orbref.Extractor(1000)
    o(img, (0, 0))
    r = refsrc.ReferenceExtractor(1000)
    for level in range(8):
        w, h = o.level_dims(level)
        cands = o.level_candidates(level)                       # minBorder-relative, octave 0, size 7, angle -1
        kept = r.distribute(cands, 16, w - 16, 16, h - 16, int(o.features_per_level[level]), level)
        want = o.level_keypoints(level)
        assert len(kept) == len(want) and len(kept) > 0
        assert np.array_equal(kept["x"] + 16, want["x"]) and np.array_equal(kept["y"] + 16, want["y"])
        assert np.array_equal(kept["response"], want["response"])


@pytest.mark.skipif(not refsrc.vocabulary_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,depth,levelsup,ragged,seed", [(10, 4, 4, True, 0), (10, 4, 2, True, 1), (6, 5, 3, True, 2),
                                                         (10, 3, 4, False, 3), (4, 6, 1, True, 4)])
def test_bow_transform_equals_dbow2_itself(tmp_path, k, depth, levelsup, ragged, seed):
    """Frame::ComputeBoW (src/Frame.cc:846-851) -> TemplatedVocabulary::transform(features, BowVector, FeatureVector,
    levelsup): DBoW2's own template (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h + FORB.cpp, compiled in place), fed a
    synthetic tree through its own loadFromTextFile, against orbref.bow_transform + the host-side accumulation the
    shim keeps (addWeight / addFeature for weight > 0, L1 normalisation)."""
    voc = synth.vocabulary(k, depth, seed, ragged)
    path = str(tmp_path / "voc.txt")
    refsrc.write_vocabulary_text(voc, k, path)
    rv = refsrc.ReferenceVocabulary(path)
    assert rv.words() == int((np.diff(voc["child_offsets"]) == 0).sum())
    rng = np.random.default_rng(seed)
    leaves = np.flatnonzero(np.diff(voc["child_offsets"]) == 0)
    # features: perturbed leaf descriptors (meaningful descents, ties at the duplicated siblings) + pure noise
    feats = np.concatenate([synth.flip_bits(voc["descriptors"][rng.choice(leaves, 600)], rng.integers(0, 40, 600), rng),
                            synth.descriptors(200, seed + 50)])
    word_r, weight_r, node_r, bow_w, bow_v = rv.transform(feats, levelsup)
    word_o, weight_o, node_o = orbref.bow_transform(orbref.make_vocabulary(**voc), feats, levelsup)
    assert np.array_equal(word_r, word_o)
    assert np.array_equal(weight_r.view(np.uint64), weight_o.view(np.uint64))
    filed = weight_o > 0                                      # stopped words are filed nowhere (:1154)
    assert filed.sum() > 500 and (~filed).sum() > 0
    # A leaf above level m_L - levelsup (possible only in a ragged tree; ORBvoc is complete) never assigns *nid, and
    # DBoW2's caller passes an uninitialised NodeId (TemplatedVocabulary.h:1149): undefined there, 0 in the oracle.
    off, children = voc["child_offsets"], voc["children"]
    depth_of = np.zeros(len(off) - 1, np.int64)
    for p in range(len(off) - 1):
        depth_of[children[off[p]:off[p + 1]]] = depth_of[p] + 1
    leaf_node = leaves[word_o.astype(np.int64)]               # word ids number the leaves in node order
    reached = depth_of[leaf_node] >= depth - levelsup
    assert np.array_equal(node_r[filed & reached], node_o[filed & reached]) and (node_r[~filed] == 0xffffffff).all()
    assert (node_o[~reached] == 0).all()
    # the BowVector: per word the sum of its features' weights in feature order, then divided by the L1 norm
    acc = {}
    for w, x in zip(word_o[filed], weight_o[filed]):
        acc[int(w)] = acc.get(int(w), 0.0) + float(x)
    norm = 0.0
    for w in sorted(acc):
        norm += abs(acc[w])
    want_w = np.array(sorted(acc), np.uint32)
    want_v = np.array([acc[w] / norm for w in sorted(acc)])
    assert np.array_equal(bow_w, want_w)
    assert np.array_equal(bow_v.view(np.uint64), want_v.view(np.uint64))



@pytest.mark.parametrize("kind,w,h,nfeat,seed", [("scene", 752, 480, 1200, 31), ("uniform_noise", 640, 480, 1000, 32)])
def test_both_keypoint_stages_of_the_reference_agree_with_the_oracle(kind, w, h, nfeat, seed):
    """The oracle follows the serial ComputeKeyPointsOctTree (:886-999); operator() calls its TBB twin (:759-885). Run
    serially the two give the same per-level keypoints (positions, responses, octave, size, angle), and those are the
    oracle's level keypoints."""
    img = synth.make(kind, h, w, seed)
    r = refsrc.ReferenceExtractor(nfeat)
    k_serial, c_serial = r.keypoints(img, True)
    k_tbb, c_tbb = r.keypoints(img, False)
    assert np.array_equal(c_serial, c_tbb) and np.array_equal(k_serial, k_tbb)
    o = orbref.Extractor(nfeat)
    o(img, (0, 0))
    at = 0
    for level in range(8):
        want = o.level_keypoints(level)
        got = k_serial[at:at + c_serial[level]]
        at += c_serial[level]
        assert len(got) == len(want)
        for f in ("x", "y", "response", "octave", "size", "angle"):
            assert np.array_equal(got[f], want[f]), (level, f)


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/src/Frame.cc"), reason="no reference tree")
def test_reference_caller_text_compiles_against_the_shim_header():
    """shim/ORBextractor.h claims header compatibility with include/ORBextractor.h for its callers. The caller on the hot
    path is Frame::ComputeStereoMatches (reads mpORBextractorLeft/Right->mvImagePyramid, src/Frame.cc:927-1029): its own
    text, cut out of the reference file, must compile with the shim's header in place of the reference's."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = subprocess.check_output(
        ["awk", r"/^void Frame::ComputeStereoMatches\(\) \{/{p=1} p{print} p&&/^}/{p=0}", "/root/reference/src/Frame.cc"],
        text=True)
    assert text.count("mvImagePyramid") >= 3
    src = ('#include <climits>\n#include "ORBmatcher.h"\n'
           "static_assert(sizeof(&ORB_SLAM3::ORBextractor::SetPyramidMirror) > 0, \"the shim header is the one in use\");\n"
           "namespace ORB_SLAM3 {\n" + text + "}\n")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-include", "oracle/ref_stubs/matcher_world.h", "-Ishim",
           "-Ioracle/ref_stubs", "-Ioracle", "-Iinclude", "-I/root/reference/include", "-I/root/reference", "-x", "c++", "-"]
    r = subprocess.run(cmd, input=src, text=True, cwd=root, capture_output=True)
    assert r.returncode == 0, r.stderr[-2000:]


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/include/ORBmatcher.h"), reason="no reference tree")
@pytest.mark.parametrize("name", ["ORBmatcher_orbx.cc", "FrameStereo_orbx.cc", "ORBmatcher_next_orbx.cc"])
def test_reference_side_shim_bodies_compile_against_the_reference_class(name):
    """shim/*.cc are the bodies a maintainer drops in place of the reference's ORBmatcher / Frame methods. They need the
    reference's headers, so they can only be type-checked where the tree exists: against the reference's own
    include/ORBmatcher.h (the class declaration they must match) over the stand-in Frame / KeyFrame / MapPoint world."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-w", "-include", "oracle/ref_stubs/matcher_world.h", "-Ishim",
           "-Ioracle/ref_stubs", "-Ioracle", "-Iinclude", "-I/root/reference/include", "-I/root/reference",
           os.path.join("shim", name)]
    r = subprocess.run(cmd, text=True, cwd=root, capture_output=True)
    assert r.returncode == 0, r.stderr[-2000:]


# What the drop-in bodies read through the stand-in classes (oracle/ref_stubs/matcher_world.h, bow_world.h) must exist,
# under the same name and with a compatible declaration, in the reference's OWN headers — otherwise a body that compiles
# and passes here would not compile in the reference tree. (ORBmatcher's methods are checked by the compiler itself: the
# stand-in world includes the reference's real include/ORBmatcher.h.) One regular expression per member.
_REAL_DECLARATIONS = [
    # shim/Tracking_orbx.cc
    ("include/Tracking.h", r"void\s+SearchLocalPoints\s*\(\s*\)\s*;"),
    ("include/Tracking.h", r"Frame\s+mCurrentFrame\s*;"),
    ("include/Tracking.h", r"std::vector<MapPoint\s*\*>\s+mvpLocalMapPoints\s*;"),
    ("include/Tracking.h", r"eTrackingState\s+mState\s*;"),
    ("include/Tracking.h", r"RECENTLY_LOST\s*=\s*3"),
    ("include/Tracking.h", r"LOST\s*=\s*4"),
    ("include/Tracking.h", r"int\s+mSensor\s*;"),
    ("include/Tracking.h", r"Atlas\s*\*\s*mpAtlas\s*;"),
    ("include/Tracking.h", r"LocalMapping\s*\*\s*mpLocalMapper\s*;"),
    ("include/Tracking.h", r"unsigned int\s+mnLastRelocFrameId\s*;"),
    ("include/System.h", r"RGBD\s*=\s*2"),
    ("include/System.h", r"IMU_MONOCULAR\s*=\s*3"),
    ("include/System.h", r"IMU_STEREO\s*=\s*4"),
    ("include/System.h", r"IMU_RGBD\s*=\s*5"),
    ("include/Atlas.h", r"bool\s+isImuInitialized\s*\(\s*\)\s*;"),
    ("include/Atlas.h", r"Map\s*\*\s*GetCurrentMap\s*\(\s*\)\s*;"),
    ("include/Map.h", r"bool\s+GetIniertialBA2\s*\(\s*\)\s*;"),
    ("include/LocalMapping.h", r"bool\s+mbFarPoints\s*;"),
    ("include/LocalMapping.h", r"float\s+mThFarPoints\s*;"),
    ("include/Frame.h", r"bool\s+isInFrustum\s*\(\s*MapPoint\s*\*\s*pMP\s*,\s*float\s+viewingCosLimit\s*\)\s*;"),
    ("include/Frame.h", r"GetPose\s*\(\s*\)\s*const"),
    ("include/Frame.h", r"Eigen::Vector3f\s+GetOw\s*\(\s*\)\s*const"),
    ("include/Frame.h", r"long unsigned int\s+mnId\s*;"),
    ("include/Frame.h", r"map<long unsigned int,\s*cv::Point2f>\s+mmProjectPoints\s*;"),
    ("include/Frame.h", r"std::vector<MapPoint\s*\*>\s+mvpMapPoints\s*;"),
    ("include/Frame.h", r"float\s+mfLogScaleFactor\s*;"),
    ("include/Frame.h", r"int\s+mnScaleLevels\s*;"),
    ("include/Frame.h", r"float\s+mbf\s*;"),
    ("include/Frame.h", r"static float\s+mnMinX\s*;"),
    ("include/Frame.h", r"GeometricCamera\s*\*\s*mpCamera\s*,"),
    ("include/Frame.h", r"int\s+Nleft\s*,"),
    ("include/CameraModels/GeometricCamera.h", r"float\s+getParameter\s*\(\s*const int i\s*\)"),
    ("include/MapPoint.h", r"void\s+IncreaseVisible\s*\(\s*int n\s*=\s*1\s*\)\s*;"),
    ("include/MapPoint.h", r"long unsigned int\s+mnLastFrameSeen\s*;"),
    ("include/MapPoint.h", r"long unsigned int\s+mnId\s*;"),
    ("include/MapPoint.h", r"bool\s+mbTrackInView\s*,\s*mbTrackInViewR\s*;"),
    ("include/MapPoint.h", r"float\s+mTrackProjX\s*;"),
    ("include/MapPoint.h", r"float\s+mTrackProjXR\s*;"),
    ("include/MapPoint.h", r"int\s+mnTrackScaleLevel\s*,"),
    ("include/MapPoint.h", r"float\s+mTrackViewCos\s*,"),
    ("include/MapPoint.h", r"float\s+mTrackDepth\s*;"),
    ("include/MapPoint.h", r"Eigen::Vector3f\s+GetWorldPos\s*\(\s*\)\s*;"),
    ("include/MapPoint.h", r"Eigen::Vector3f\s+GetNormal\s*\(\s*\)\s*;"),
    ("include/MapPoint.h", r"float\s+mfMinDistance\s*;"),      # what the two accessors to add return
    ("include/MapPoint.h", r"float\s+mfMaxDistance\s*;"),
    # shim/FrameBoW_orbx.cc
    ("include/Frame.h", r"void\s+ComputeBoW\s*\(\s*\)\s*;"),
    ("include/KeyFrame.h", r"void\s+ComputeBoW\s*\(\s*\)\s*;"),
    ("include/Frame.h", r"ORBVocabulary\s*\*\s*mpORBvocabulary\s*;"),
    ("include/KeyFrame.h", r"ORBVocabulary\s*\*\s*mpORBvocabulary\s*;"),
    ("include/Frame.h", r"DBoW2::BowVector\s+mBowVec\s*;"),
    ("include/Frame.h", r"DBoW2::FeatureVector\s+mFeatVec\s*;"),
    ("include/KeyFrame.h", r"DBoW2::BowVector\s+mBowVec\s*;"),
    ("include/KeyFrame.h", r"DBoW2::FeatureVector\s+mFeatVec\s*;"),
    ("include/Frame.h", r"cv::Mat\s+mDescriptors\s*,"),
    ("include/KeyFrame.h", r"const cv::Mat\s+mDescriptors\s*;"),
    ("Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h", r"std::vector<Node>\s+m_nodes\s*;"),
    ("Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h", r"GeneralScoring\s*\*\s*m_scoring_object\s*;"),
    ("Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h", r"inline int getDepthLevels\s*\(\s*\)\s*const"),
    ("Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h", r"inline WeightingType getWeightingType\s*\(\s*\)\s*const"),
    # shim/FrameStereo_orbx.cc, shim/ORBmatcher_next_orbx.cc
    ("include/Frame.h", r"void\s+ComputeStereoMatches\s*\(\s*\)\s*;"),
    ("include/Frame.h", r"void\s+ComputeStereoFishEyeMatches\s*\(\s*\)\s*;"),
    ("include/Frame.h", r"void\s+AssignFeaturesToGrid\s*\(\s*\)\s*;"),
    ("include/MapPoint.h", r"void\s+ComputeDistinctiveDescriptors\s*\(\s*\)\s*;"),
]


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/include/Tracking.h"), reason="no reference tree")
def test_members_the_drop_in_bodies_use_exist_in_the_reference_headers():
    import re
    cache, missing = {}, []
    for rel, pattern in _REAL_DECLARATIONS:
        if rel not in cache:
            cache[rel] = open("/root/reference/" + rel, errors="replace").read()
        if not re.search(pattern, cache[rel]):
            missing.append((rel, pattern))
    assert not missing, missing


@pytest.mark.skipif(not __import__("os").path.exists("/root/reference/include/Tracking.h"), reason="no reference tree")
def test_every_member_name_the_shim_reaches_is_a_name_of_the_reference():
    """Word-level net under the list above: every identifier the files of shim/ reach with `.` or `->` must occur in the
    reference's headers (include/, CameraModels/, DBoW2), in this repository's C ABI headers, or be a standard-library /
    OpenCV / Eigen method — except the accessors INTEGRATION.md asks a maintainer to add, which must be named there."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = "".join(open(f, errors="replace").read() for pat in ("include/*.h", "include/CameraModels/*.h", "Thirdparty/DBoW2/DBoW2/*.h")
                  for f in glob.glob("/root/reference/" + pat))
    own = "".join(open(f).read() for f in glob.glob(os.path.join(root, "include", "*.h")))
    known = set(re.findall(r"[A-Za-z_]\w*", ref)) | set(re.findall(r"[A-Za-z_]\w*", own))
    library = {"getMat", "release", "dot"}                      # OpenCV / Eigen methods no reference header spells out
    to_add = {"GetGridCell", "GetGridCellRight", "GetMinDistanceRaw", "GetMaxDistanceRaw"}   # accessors a maintainer adds
    shim_own = {"Handle", "has_mp"}               # shim/ORBextractor.h's own accessor; a field of a struct local to a shim file
    integration = open(os.path.join(root, "INTEGRATION.md")).read()
    for name in to_add:
        assert name in integration, name
    for path in sorted(glob.glob(os.path.join(root, "shim", "*"))):
        text = re.sub(r"//.*", "", open(path).read())
        reached = set(re.findall(r"(?:->|\.)\s*([A-Za-z_]\w*)", text))
        unknown = sorted(n for n in reached if n not in known and n not in library and n not in to_add and n not in shim_own)
        assert not unknown, (os.path.basename(path), unknown)
