"""Parity of the CUDA extractor (through the C ABI) with the CPU oracle, stage by stage and end to end.

Bar (BASELINE.json north_star): bit-exact keypoint coordinates / octave / descriptors; angle and response within 1e-4
(they are in fact compared bit-exactly here, the tolerance is only the documented bar).
"""
import numpy as np
import pytest

from orb_slam3_fast_b200 import ORBextractor, OrbxError, synth
from oracle import orbref

pytestmark = pytest.mark.gpu

ANGLE_TOL = 1e-4


def _compare_frame(ex_ref, kps, desc, mono, img, lap, tag):
    mono_r, kps_r, desc_r = ex_ref(img, lap)
    assert len(kps) == len(kps_r), "%s: keypoint count %d vs oracle %d" % (tag, len(kps), len(kps_r))
    assert mono == mono_r, "%s: monoIndex %d vs %d" % (tag, mono, mono_r)
    for fld in ("x", "y", "size", "octave", "class_id", "response"):
        bad = np.nonzero(kps[fld] != kps_r[fld])[0]
        assert len(bad) == 0, "%s: field %s differs at %s" % (tag, fld, bad[:10])
    da = np.abs(kps["angle"] - kps_r["angle"])
    assert (da <= ANGLE_TOL).all(), "%s: angle differs by up to %g" % (tag, da.max())
    assert (kps["angle"] == kps_r["angle"]).all(), "%s: angle not bit-exact (max diff %g)" % (tag, da.max())
    bad = np.nonzero((desc != desc_r).any(axis=1))[0]
    assert len(bad) == 0, "%s: %d descriptors differ, first rows %s" % (tag, len(bad), bad[:10])


def _stage_check(ex, ex_ref, tag):
    for l in range(ex.nlevels):
        assert ex.level_size(l) == ex_ref.level_dims(l), "%s: level %d size" % (tag, l)
        raw, raw_r = ex.debug_level(l), ex_ref.level_image(l)
        assert np.array_equal(raw, raw_r), "%s: pyramid level %d differs in %d px" % (tag, l, (raw != raw_r).sum())
    for l in range(ex.nlevels):
        c, c_r = ex.debug_candidates(l), ex_ref.level_candidates(l)
        assert len(c) == len(c_r), "%s: level %d candidates %d vs %d" % (tag, l, len(c), len(c_r))
        for fld in ("x", "y", "response"):
            assert np.array_equal(c[fld], c_r[fld]), "%s: level %d candidate %s" % (tag, l, fld)
    for l in range(ex.nlevels):
        k, k_r = ex.debug_level_keypoints(l), ex_ref.level_keypoints(l)
        assert len(k) == len(k_r), "%s: level %d quadtree count %d vs %d" % (tag, l, len(k), len(k_r))
        for fld in ("x", "y", "response", "size", "octave"):
            assert np.array_equal(k[fld], k_r[fld]), "%s: level %d quadtree %s" % (tag, l, fld)
    for l in range(ex.nlevels):
        b_r = ex_ref.level_blurred(l)
        if b_r is None:
            continue
        b = ex.debug_level(l, blurred=True)
        assert np.array_equal(b, b_r), "%s: blurred level %d differs in %d px" % (tag, l, (b != b_r).sum())


CASES = [
    ("scene", 480, 640, 1000, (0, 0)),
    ("scene", 480, 640, 1000, (0, 1000)),   # monocular lapping: everything written from the back
    ("scene", 480, 752, 1200, (0, 0)),
    ("noise_blur", 480, 640, 1000, (0, 0)),
    ("uniform_noise", 480, 640, 1000, (0, 0)),
    ("scene", 720, 1280, 2000, (0, 1000)),
    ("scene", 241, 241, 500, (0, 0)),        # minimum size for 8 levels
    ("scene", 300, 900, 300, (100, 500)),    # 3 root nodes, partial lapping
]


@pytest.mark.parametrize("kind,h,w,nf,lap", CASES)
def test_extract_matches_oracle(gpu, kind, h, w, nf, lap):
    img = synth.make(kind, h, w, seed=3)
    ex = ORBextractor(nf)
    ex_ref = orbref.Extractor(nf)
    mono, kps, desc = ex(img, lap)
    tag = "%s %dx%d nf=%d" % (kind, w, h, nf)
    mono_r, kps_r, desc_r = ex_ref(img, lap)
    _stage_check(ex, ex_ref, tag)
    _compare_frame(ex_ref, kps, desc, mono, img, lap, tag)


def test_low_contrast_uses_min_threshold(gpu):
    img = synth.low_contrast(480, 640, seed=1)
    ex, ex_ref = ORBextractor(1000), orbref.Extractor(1000)
    mono, kps, desc = ex(img)
    ex_ref(img, (0, 0))
    _stage_check(ex, ex_ref, "low_contrast")
    _compare_frame(ex_ref, kps, desc, mono, img, (0, 0), "low_contrast")


def test_constant_image_gives_no_keypoints(gpu):
    ex = ORBextractor(1000)
    mono, kps, desc = ex(synth.constant(480, 640))
    assert mono == 0 and len(kps) == 0 and desc.shape == (0, 32)


def test_empty_image_returns_minus_one(gpu):
    ex = ORBextractor(1000)
    mono, kps, desc = ex(np.empty((0, 0), np.uint8))
    assert mono == -1 and len(kps) == 0


def test_too_small_image_is_an_error(gpu):
    ex = ORBextractor(1000)
    with pytest.raises(OrbxError) as e:
        ex(synth.scene(100, 100, 0))
    assert e.value.code == -5


def test_strided_input_and_size_change(gpu):
    big = synth.scene(500, 700, seed=5)
    view = big[10:490, 20:660]  # non-contiguous rows
    ex, ex_ref = ORBextractor(1000), orbref.Extractor(1000)
    mono, kps, desc = ex(view)
    _compare_frame(ex_ref, kps, desc, mono, np.ascontiguousarray(view), (0, 0), "strided")
    img2 = synth.scene(480, 752, seed=6)  # same handle, new geometry
    mono, kps, desc = ex(img2)
    _compare_frame(ex_ref, kps, desc, mono, img2, (0, 0), "resized")


def test_batch_matches_single_and_oracle(gpu):
    imgs = np.stack([synth.make(k, 480, 640, s) for s, k in enumerate(["scene", "noise_blur", "scene", "uniform_noise",
                                                                        "scene"])])
    ex = ORBextractor(1000, max_batch=3)  # 5 frames through groups of 3 + 2
    ex_ref = orbref.Extractor(1000)
    n, mono, kps, desc = ex.extract_batch(imgs, (0, 0))
    for f in range(len(imgs)):
        _compare_frame(ex_ref, kps[f, :n[f]], desc[f, :n[f]], mono[f], imgs[f], (0, 0), "batch frame %d" % f)


def test_frame_index_addresses_the_whole_call_not_the_last_group(gpu):
    """`frame` of orbx_download_pyramid / orbx_debug_* / orbm_stereo_match is the index inside the last call: with
    n_frames > max_batch the frames of earlier groups stay addressable while their lane is resident (the last 8 groups),
    and a frame that is gone is an error instead of another frame's data."""
    from orb_slam3_fast_b200 import ORBmatcher
    from orb_slam3_fast_b200.lib import OrbxError
    n = 19
    imgs = np.stack([synth.stereo_pair(480, 640, 70 + s)[0] for s in range(n)])
    right = np.stack([synth.stereo_pair(480, 640, 70 + s)[1] for s in range(n)])
    ex, exr = ORBextractor(1000, max_batch=2), ORBextractor(1000, max_batch=2)  # 10 groups of 2 (the last of 1) on 8 lanes: groups 0 and 1 are overwritten
    ex_ref, exr_ref = orbref.Extractor(1000), orbref.Extractor(1000)
    nk, mono, kps, desc = ex.extract_batch(imgs, (0, 0))
    nr, _, kr, dr = exr.extract_batch(right, (0, 0))
    mt = ORBmatcher()
    mbf, mb = float(np.float32(435.2 * 0.11)), float(np.float32(0.11))
    for f in (4, 7, 16, 18):  # resident: lanes 2, 3, 0, 1
        ex_ref(imgs[f], (0, 0))
        exr_ref(right[f], (0, 0))
        for l in (0, 5):
            assert np.array_equal(ex.image_pyramid_bordered(l, f), ex_ref.level_bordered(l)), "frame %d level %d" % (f, l)
            assert np.array_equal(ex.debug_level(l, True, f), ex_ref.level_blurred(l))
            assert np.array_equal(ex.debug_level_keypoints(l, f)[["x", "y"]], ex_ref.level_keypoints(l)[["x", "y"]])
        got = mt.ComputeStereoMatches(ex, exr, kps[f, :nk[f]], desc[f, :nk[f]], kr[f, :nr[f]], dr[f, :nr[f]], mbf, mb, frame=f)
        ref = orbref.stereo_match(ex_ref, exr_ref, kps[f, :nk[f]], desc[f, :nk[f]], kr[f, :nr[f]], dr[f, :nr[f]], mbf, mb)
        assert got[0] == ref[0] and got[1].tobytes() == ref[1].tobytes() and got[2].tobytes() == ref[2].tobytes(), f
    for f in (0, 3, 19, -1):  # overwritten by the second round of groups / out of range
        with pytest.raises(OrbxError):
            ex.image_pyramid_bordered(0, f)
    ex(imgs[0])  # a new call: the old indices are gone
    with pytest.raises(OrbxError):
        ex.image_pyramid_bordered(0, 16)
    assert np.array_equal(ex.image_pyramid(0, 0), imgs[0])


def test_host_pyramid_mirror_has_reference_border(gpu):
    img = synth.scene(480, 640, seed=9)
    ex, ex_ref = ORBextractor(1000), orbref.Extractor(1000)
    ex(img)
    ex_ref(img, (0, 0))
    for l in (0, 3, 7):
        assert np.array_equal(ex.image_pyramid_bordered(l), ex_ref.level_bordered(l)), "level %d" % l


def test_other_parameters(gpu):
    img = synth.scene(480, 640, seed=11)
    # (…, 1.1, 10, …): more than 8 levels -> the LDG kernels (the tensor-map arrays hold 8); (…, 2.0, 3, …): a source
    # rectangle wider than one TMA box -> the LDG resize
    for nf, sf, nl, it, mt in ((500, 1.2, 4, 20, 7), (1500, 1.1, 8, 15, 5), (5000, 1.2, 8, 20, 7), (300, 1.5, 3, 30, 10),
                               (800, 1.1, 10, 20, 7), (300, 2.0, 3, 20, 7)):
        ex = ORBextractor(nf, sf, nl, it, mt)
        ex_ref = orbref.Extractor(nf, sf, nl, it, mt)
        assert np.array_equal(ex.mnFeaturesPerLevel, ex_ref.features_per_level)
        assert np.array_equal(ex.GetScaleFactors(), ex_ref.scale)
        mono, kps, desc = ex(img)
        _compare_frame(ex_ref, kps, desc, mono, img, (0, 0), "params %s" % ((nf, sf, nl, it, mt),))


@pytest.mark.parametrize("base_off,pad", [(0, 0), (0, 16), (0, 4), (0, 3), (1, 0), (4, 8)])
def test_device_resident_input_every_staging_variant(gpu, base_off, pad):
    """orbx_extract_batch_device reads level 0 in place from the caller's device buffer. A 16-byte aligned base / pitch
    / frame stride takes the TMA kernels (k_fast<true>, k_resize_tma, k_blur7<kBlurTma>); anything else must fall back
    to the LDG kernels (word aligned: k_blur7<kBlurLdg>; odd: byte loads) — all of them bit-exact with the oracle."""
    import torch
    from orb_slam3_fast_b200.synth import KP_DTYPE
    h, w, nf, B = 480, 640, 1000, 3
    imgs = [synth.scene(h, w, 40 + k) for k in range(B)]
    pitch = w + pad
    fstride = pitch * h + (0 if (base_off == 0 and pad % 16 == 0) else 0)
    buf = torch.zeros(base_off + B * fstride + 64, dtype=torch.uint8, device="cuda")
    for k, im in enumerate(imgs):
        view = buf[base_off + k * fstride: base_off + (k + 1) * fstride].view(h, pitch)
        view[:, :w] = torch.from_numpy(im).cuda()
        if pad:
            view[:, w:] = 255 - view[:, :pad]   # the padding is not image data: it must never influence the result
    ex = ORBextractor(nf, max_batch=B)
    cap = ex.capacity
    d_kps = torch.empty((B, cap, 7), dtype=torch.int32, device="cuda")
    d_desc = torch.empty((B, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.empty(B, dtype=torch.int32, device="cuda")
    d_mono = torch.empty(B, dtype=torch.int32, device="cuda")
    d_status = torch.empty(B, dtype=torch.int32, device="cuda")
    ex.extract_batch_device(buf.data_ptr() + base_off, B, w, h, pitch, fstride, (0, 0), d_kps.data_ptr(),
                            d_desc.data_ptr(), cap, d_n.data_ptr(), d_mono.data_ptr(), d_status.data_ptr())
    torch.cuda.synchronize()
    assert (d_status.cpu().numpy() == 0).all()
    n = d_n.cpu().numpy()
    kps = d_kps.cpu().numpy().view(np.uint8).reshape(B, cap, 28).copy().view(KP_DTYPE).reshape(B, cap)
    desc = d_desc.cpu().numpy()
    ref = orbref.Extractor(nf)
    for k in range(B):
        _compare_frame(ref, kps[k, :n[k]], desc[k, :n[k]], int(d_mono[k].item()), imgs[k], (0, 0),
                       "device input base+%d pitch+%d frame %d" % (base_off, pad, k))


def test_blur_kernels_tensor_core_form_and_fallback(gpu):
    """The Gaussian blur runs as banded u8 GEMMs on the tensor cores (k_blur_tc) whenever TMA can address the levels; the
    CUDA-core kernel k_blur7<kBlurTma> is then only reached with ORBX_BLUR_TC=0 (read once per process, hence the
    subprocess). Both must give the oracle's keypoints and descriptors — the descriptor bits are sign tests on the blurred
    level, so a single wrong byte shows — on sizes whose levels have one tile column / row, ragged right / bottom tiles
    and widths that are not multiples of 16."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from orb_slam3_fast_b200 import ORBextractor, synth
from oracle import orbref
for (h, w, nf, nl) in ((480, 640, 1000, 8), (241, 333, 500, 5), (130, 517, 400, 3), (720, 1280, 1500, 8)):
    img = synth.scene(h, w, seed=h + w)
    ex, ref = ORBextractor(nf, 1.2, nl), orbref.Extractor(nf, 1.2, nl)
    m, k, d = ex(img, (0, 0))
    m_r, k_r, d_r = ref(img, (0, 0))
    assert m == m_r and np.array_equal(k, k_r) and np.array_equal(d, d_r), (h, w)
    imgs = np.stack([synth.scene(h, w, seed=h + w + s) for s in range(3)])
    exb = ORBextractor(nf, 1.2, nl, max_batch=3)
    n_out, _, kps, desc = exb.extract_batch(imgs)
    for s in range(3):
        _, k_r, d_r = ref(imgs[s], (0, 0))
        n = int(n_out[s])
        assert n == len(k_r) and np.array_equal(kps[s, :n], k_r) and np.array_equal(desc[s, :n], d_r), (h, w, s)
print("ok")
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for tc in ("1", "0"):
        env = dict(os.environ, ORBX_BLUR_TC=tc)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "ok" in r.stdout, "ORBX_BLUR_TC=%s\n%s\n%s" % (tc, r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("channels,rgb", [(3, False), (3, True), (4, False), (4, True)])
def test_cvt_color_to_gray(gpu, channels, rgb):
    """cv::cvtColor(..., COLOR_*2GRAY) (src/Tracking.cc:1394-1412) on the device == oracle (== cv2, CPU suite); widths
    that are not a multiple of 4 and the batched device form on a strided buffer."""
    import torch
    from orb_slam3_fast_b200 import cvtColorToGray
    from orb_slam3_fast_b200 import lib as _lib
    rng = np.random.default_rng(channels * 2 + int(rgb))
    for (h, w) in ((480, 752), (33, 101), (5, 3)):
        img = rng.integers(0, 256, (h, w, channels), dtype=np.uint8)
        got = cvtColorToGray(img, rgb)
        assert np.array_equal(got, orbref.cvt_gray(img, rgb)), (h, w)
        import cv2  # the real OpenCV kernel, directly
        code = {(3, False): cv2.COLOR_BGR2GRAY, (3, True): cv2.COLOR_RGB2GRAY, (4, False): cv2.COLOR_BGRA2GRAY,
                (4, True): cv2.COLOR_RGBA2GRAY}[(channels, rgb)]
        assert np.array_equal(got, cv2.cvtColor(img, code)), (h, w)
    # device form: 3 frames, padded rows on both sides, converted straight into an extractor-ready gray batch
    B, h, w, spad, dpad = 3, 120, 200, 8 * channels, 16
    imgs = rng.integers(0, 256, (B, h, w + 8, channels), dtype=np.uint8)
    d_src = torch.from_numpy(imgs).cuda()
    d_dst = torch.zeros((B, h, w + dpad), dtype=torch.uint8, device="cuda")
    rc = _lib.lib().orbx_cvt_gray_device(0, B, d_src.data_ptr(), w, h, (w + 8) * channels, h * (w + 8) * channels, channels,
                                         int(rgb), d_dst.data_ptr(), w + dpad, h * (w + dpad), None)
    assert rc == 0
    torch.cuda.synchronize()
    out = d_dst.cpu().numpy()
    for k in range(B):
        assert np.array_equal(out[k, :, :w], orbref.cvt_gray(np.ascontiguousarray(imgs[k, :, :w]), rgb))
    assert (out[:, :, w:] == 0).all()


def test_remap_linear_rectification(gpu):
    """cv::remap(..., INTER_LINEAR) (src/System.cc:293-294) on the device == oracle (== cv2, CPU suite): host form on two
    geometries, device form on a batch that shares one pair of maps."""
    import torch
    from orb_slam3_fast_b200 import remapLinear
    from orb_slam3_fast_b200 import lib as _lib
    from test_oracle_primitives import _rectify_maps
    rng = np.random.default_rng(12)
    for (h, w, dh, dw) in ((480, 752, 480, 752), (120, 161, 97, 203)):
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        mapx, mapy = _rectify_maps(h, w, h + w, dh, dw)
        got = remapLinear(src, mapx, mapy)
        assert np.array_equal(got, orbref.remap_linear(src, mapx, mapy)), (h, w)
        import cv2  # the real OpenCV kernel, directly
        assert np.array_equal(got, cv2.remap(src, mapx, mapy, cv2.INTER_LINEAR)), (h, w)
    B, h, w = 3, 240, 320
    imgs = rng.integers(0, 256, (B, h, w), dtype=np.uint8)
    mapx, mapy = _rectify_maps(h, w, 5)
    d_src, d_mx, d_my = torch.from_numpy(imgs).cuda(), torch.from_numpy(mapx).cuda(), torch.from_numpy(mapy).cuda()
    d_dst = torch.zeros((B, h, w), dtype=torch.uint8, device="cuda")
    rc = _lib.lib().orbx_remap_linear_device(0, B, d_src.data_ptr(), w, h, w, h * w, d_mx.data_ptr(), d_my.data_ptr(), w, h,
                                             d_dst.data_ptr(), w, h * w, None)
    assert rc == 0
    torch.cuda.synchronize()
    out = d_dst.cpu().numpy()
    for k in range(B):
        assert np.array_equal(out[k], orbref.remap_linear(imgs[k], mapx, mapy))
