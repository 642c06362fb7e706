"""Second, independent restatement of ORBextractor::operator() that drives the REAL OpenCV kernels through cv2.

TEST INFRASTRUCTURE ONLY (same rules as orbref: tests/ and golden generation may import it, the product never does).

Why it exists: the reference cannot be built in this image (no OpenCV C++ / TBB / Eigen), so the C++ oracle
(oracle/orbref.cpp) restates both the reference's orchestration AND OpenCV's arithmetic. This module restates only
the orchestration (src/ORBextractor.cc, serial path) in Python and calls the genuine OpenCV 4.13 implementations of
cv::resize, cv::copyMakeBorder, cv::FAST, cv::GaussianBlur and cv::fastAtan2, glibc cosf/sinf through ctypes and
libstdc++ std::sort through orbref.std_sort_perm. tests/test_oracle_pipeline.py requires the two to agree bit for
bit; tests/golden/*.npz are frozen from this module by tools/make_golden.py.
"""
import ctypes
import math

import cv2
import numpy as np

from . import orbref

_libm = ctypes.CDLL("libm.so.6")
_libm.cosf.argtypes = [ctypes.c_float]
_libm.cosf.restype = ctypes.c_float
_libm.sinf.argtypes = [ctypes.c_float]
_libm.sinf.restype = ctypes.c_float
_libm.lrintf.argtypes = [ctypes.c_float]
_libm.lrintf.restype = ctypes.c_long

EDGE = 19
F32 = np.float32
_PATTERN = None


def _pattern():
    global _PATTERN
    if _PATTERN is None:
        import os
        import re
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "orb_slam3_fast_b200", "csrc",
                         "orb_pattern.inc")
        txt = "\n".join(l for l in open(p) if not l.lstrip().startswith("//"))
        _PATTERN = np.array([int(v) for v in re.findall(r"-?\d+", txt)], np.int32).reshape(256, 4)
    return _PATTERN


def cv_round(v):
    return int(_libm.lrintf(float(F32(v))))


class Node:
    __slots__ = ("keys", "ulx", "uly", "urx", "bly", "leaf", "alive")

    def __init__(self, ulx, uly, urx, bly):
        self.keys, self.ulx, self.uly, self.urx, self.bly, self.leaf, self.alive = [], ulx, uly, urx, bly, False, True


def _split(p, cx, cy):
    halfx = int(math.ceil(float(F32(p.urx - p.ulx) / F32(2))))
    halfy = int(math.ceil(float(F32(p.bly - p.uly) / F32(2))))
    mx, my = p.ulx + halfx, p.uly + halfy
    ch = [Node(p.ulx, p.uly, mx, my), Node(mx, p.uly, p.urx, my), Node(p.ulx, my, mx, p.bly),
          Node(mx, my, p.urx, p.bly)]
    for k in p.keys:
        q = (0 if cx[k] < mx else 1) + (0 if cy[k] < my else 2)
        ch[q].keys.append(k)
    for c in ch:
        c.leaf = len(c.keys) == 1
    return ch


def distribute(cx, cy, resp, min_x, max_x, min_y, max_y, n_want):
    """DistributeOctTree (src/ORBextractor.cc:557-757) on candidate arrays; returns the selected candidate indices in
    list order. The std::list is a Python list with index 0 = front."""
    n_ini = int(round(float(F32(max_x - min_x) / F32(max_y - min_y))))
    hx = F32(max_x - min_x) / F32(n_ini)
    nodes = [Node(int(hx * F32(i)), 0, int(hx * F32(i + 1)), max_y - min_y) for i in range(n_ini)]
    for k in range(len(cx)):
        nodes[int(F32(cx[k]) / hx)].keys.append(k)
    nodes = [n for n in nodes if n.keys]
    for n in nodes:
        n.leaf = len(n.keys) == 1
    done = False
    while not done:
        prev = len(nodes)
        pending, front, n_expand = [], [], 0
        keep = []
        for n in nodes:
            if n.leaf:
                keep.append(n)
                continue
            for c in _split(n, cx, cy):
                if c.keys:
                    front.insert(0, c)
                    if len(c.keys) > 1:
                        n_expand += 1
                        pending.append(c)
        nodes = front + keep
        if len(nodes) >= n_want or len(nodes) == prev:
            done = True
        elif len(nodes) + 3 * n_expand > n_want:
            while not done:
                prev = len(nodes)
                work, pending = pending, []
                perm = orbref.std_sort_perm([len(n.keys) for n in work], [n.ulx for n in work])
                work = [work[i] for i in perm]
                for j in range(len(work) - 1, -1, -1):
                    for c in _split(work[j], cx, cy):
                        if c.keys:
                            nodes.insert(0, c)
                            if len(c.keys) > 1:
                                pending.append(c)
                    nodes.remove(work[j])
                    if len(nodes) >= n_want:
                        break
                if len(nodes) >= n_want or len(nodes) == prev:
                    done = True
    out = []
    for n in nodes:
        best = n.keys[0]
        for k in n.keys[1:]:
            if resp[k] > resp[best]:
                best = k
        out.append(best)
    return out


class Cv2Extractor:
    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sfd = float(F32(scale_factor))  # the double member initialised from a float
        self.sf = [F32(1.0)]
        for _ in range(1, nlevels):
            self.sf.append(F32(float(self.sf[-1]) * sfd))
        self.inv_sf = [F32(1.0) / s for s in self.sf]
        factor = F32(1.0 / sfd)
        want = F32(nfeatures) * (F32(1) - factor) / (F32(1) - F32(math.pow(float(factor), float(nlevels))))
        self.per_level, total = [], 0
        for _ in range(nlevels - 1):
            self.per_level.append(cv_round(want))
            total += self.per_level[-1]
            want = want * factor
        self.per_level.append(max(nfeatures - total, 0))
        self.umax = [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
        self.pyramid = []

    def __call__(self, img, lapping=(0, 0)):
        h0, w0 = img.shape
        pyr = []
        for l in range(self.nlevels):
            w, h = cv_round(F32(w0) * self.inv_sf[l]), cv_round(F32(h0) * self.inv_sf[l])
            lvl = img if l == 0 else cv2.resize(pyr[-1], (w, h), interpolation=cv2.INTER_LINEAR)
            pyr.append(np.ascontiguousarray(lvl))
        self.pyramid = pyr
        fast = {t: cv2.FastFeatureDetector_create(t, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
                for t in (self.ini_th, self.min_th)}
        levels = []
        for l, lvl in enumerate(pyr):
            h, w = lvl.shape
            min_bx = min_by = EDGE - 3
            max_bx, max_by = w - EDGE + 3, h - EDGE + 3
            width, height = F32(max_bx - min_bx), F32(max_by - min_by)
            n_cols, n_rows = int(width / F32(35)), int(height / F32(35))
            w_cell, h_cell = int(math.ceil(float(width / F32(n_cols)))), int(math.ceil(float(height / F32(n_rows))))
            cx, cy, resp = [], [], []
            for i in range(n_rows):
                ini_y = min_by + i * h_cell
                max_y = ini_y + h_cell + 6
                if ini_y >= max_by - 3:
                    continue
                max_y = min(max_y, max_by)
                for j in range(n_cols):
                    ini_x = min_bx + j * w_cell
                    max_x = ini_x + w_cell + 6
                    if ini_x >= max_bx - 6:
                        continue
                    max_x = min(max_x, max_bx)
                    cell = np.ascontiguousarray(lvl[ini_y:max_y, ini_x:max_x])
                    kps = fast[self.ini_th].detect(cell)
                    if not kps:
                        kps = fast[self.min_th].detect(cell)
                    for k in kps:
                        cx.append(F32(k.pt[0]) + F32(j * w_cell))
                        cy.append(F32(k.pt[1]) + F32(i * h_cell))
                        resp.append(F32(k.response))
            sel = distribute(cx, cy, resp, min_bx, max_bx, min_by, max_by, self.per_level[l]) if cx else []
            size = F32(int(F32(31) * self.sf[l]))
            kl = []
            for k in sel:
                x, y = cx[k] + F32(min_bx), cy[k] + F32(min_by)
                kl.append([x, y, size, self._angle(lvl, x, y), resp[k], l])
            levels.append(kl)
        total = sum(len(k) for k in levels)
        kps = np.zeros(total, orbref.KP_DTYPE)
        desc = np.zeros((total, 32), np.uint8)
        mono, stereo = 0, total - 1
        self.blurred = []
        for l, kl in enumerate(levels):
            if not kl:
                self.blurred.append(None)
                continue
            blur = cv2.GaussianBlur(pyr[l].copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
            self.blurred.append(blur)
            for x, y, size, ang, resp_k, octv in kl:
                d = self._descriptor(blur, x, y, ang)
                if l != 0:
                    x, y = x * self.sf[l], y * self.sf[l]
                if F32(lapping[0]) <= x <= F32(lapping[1]):
                    dst = stereo
                    stereo -= 1
                else:
                    dst = mono
                    mono += 1
                kps[dst] = (x, y, size, ang, resp_k, octv, -1)
                desc[dst] = d
        return mono, kps, desc

    def _angle(self, lvl, x, y):
        cxp, cyp = cv_round(x), cv_round(y)
        m01 = m10 = 0
        row = lvl[cyp].astype(np.int64)
        for u in range(-15, 16):
            m10 += u * int(row[cxp + u])
        for v in range(1, 16):
            d = self.umax[v]
            a = lvl[cyp + v, cxp - d:cxp + d + 1].astype(np.int64)
            b = lvl[cyp - v, cxp - d:cxp + d + 1].astype(np.int64)
            u = np.arange(-d, d + 1)
            m01 += v * int((a - b).sum())
            m10 += int((u * (a + b)).sum())
        return F32(cv2.fastAtan2(float(m01), float(m10)))

    def _descriptor(self, blur, x, y, ang):
        rad = F32(ang) * F32(math.pi / float(F32(180.0)))
        a, b = F32(_libm.cosf(float(rad))), F32(_libm.sinf(float(rad)))
        cxp, cyp = cv_round(x), cv_round(y)
        pat = _pattern()
        out = np.zeros(32, np.uint8)
        for i in range(256):
            x0, y0, x1, y1 = (F32(v) for v in pat[i])
            t0 = blur[cyp + cv_round(x0 * b + y0 * a), cxp + cv_round(x0 * a - y0 * b)]
            t1 = blur[cyp + cv_round(x1 * b + y1 * a), cxp + cv_round(x1 * a - y1 * b)]
            if t0 < t1:
                out[i >> 3] |= 1 << (i & 7)
        return out
