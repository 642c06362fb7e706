"""The C-ABI shared library without a GPU: it loads, exports every symbol include/*.h declares, mirrors the reference's
error behaviour at the boundary, and FAILS LOUDLY (ORBX_E_CUDA) instead of computing anything on the host."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = []
    for hdr in ("orbx.h", "orbm.h"):
        txt = open(os.path.join(INCLUDE, hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(orb[xm]_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_headers_declare_the_expected_surface():
    names = declared_symbols()
    for must in ("orbx_extractor_create", "orbx_extract", "orbx_extract_batch", "orbx_extract_batch_device",
                 "orbx_download_pyramid", "orbx_extractor_tables", "orbm_knn2", "orbm_stereo_match",
                 "orbm_stereo_frames_batch", "orbm_search_by_projection_map", "orbm_search_by_projection_frame",
                 "orbm_search_for_triangulation", "orbm_descriptor_distance_batch"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from orb_slam3_fast_b200 import lib
    L = lib.lib()
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, "declared in include/*.h but not exported by liborbx.so: %s" % missing


def test_headers_compile_as_plain_c(tmp_path):
    src = tmp_path / "abi.c"
    src.write_text('#include "orbx.h"\n#include "orbm.h"\n'
                   "int main(void) { orbx_kp k; orbx_frame_view f; (void)k; (void)f; "
                   "return sizeof(orbx_kp) == 28 ? 0 : 1; }\n")
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", INCLUDE, str(src), "-o", str(exe)])
    assert subprocess.call([str(exe)]) == 0   # cv::KeyPoint is 28 bytes (include/Frame.h:254)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from orb_slam3_fast_b200 import ORBextractor, ORBmatcher, OrbxError
    with pytest.raises(OrbxError) as e:
        ORBextractor(1000)
    assert e.value.code == -4 and "CUDA" in str(e.value)
    with pytest.raises(OrbxError) as e:
        ORBmatcher()
    assert e.value.code == -4


def test_argument_errors_are_codes_not_crashes():
    from orb_slam3_fast_b200 import lib
    L = lib.lib()
    h = C.c_void_p()
    assert L.orbx_extractor_create(None, 0, 1000, C.c_float(1.2), 8, 20, 7, 1) == -3
    assert L.orbx_extractor_create(C.byref(h), 0, 0, C.c_float(1.2), 8, 20, 7, 1) == -3       # nfeatures < 1
    assert L.orbx_extractor_create(C.byref(h), 0, 1000, C.c_float(1.0), 8, 20, 7, 1) == -3    # scale <= 1
    assert L.orbx_extractor_create(C.byref(h), 0, 1000, C.c_float(1.2), 99, 20, 7, 1) == -3   # too many levels
    assert h.value is None
    assert L.orbx_extractor_levels(None) == -3
    assert L.orbx_extract(None, None, 0, 0, 0, 0, 0, None, None, 0, None, None) == -3
    assert L.orbx_last_error(None) is not None
    L.orbx_extractor_destroy(None)   # no-op
    assert L.orbx_host_alloc(0) is None


def test_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orb_slam3_fast_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".inc")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "orbref" not in txt and "oracle/" not in txt and "from oracle" not in txt, \
                    "%s references the oracle" % os.path.join(dirpath, f)
    out = subprocess.check_output(["ldd", os.path.join(pkg, "liborbx.so")], text=True)
    assert "orbref" not in out


def test_every_entry_point_answers_null_arguments_with_a_code():
    """"Never throw, never exit, never crash" (include/orbx.h:8-11): every int-returning entry point of both headers,
    called with a NULL handle, NULL pointers and zero sizes, must come back with a negative error code. The calls are
    generated from the declarations and run in a child process, so that a crash is a test failure naming the entry point
    (orbm_search_by_bow / _kf used to read their views before checking them)."""
    import re
    import sys
    decls = []
    for h in ("orbx.h", "orbm.h"):
        text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", h)).read(), flags=re.S)
        for m in re.finditer(r"\bint\s+(orb[xm]_\w+)\s*\(([^;]*?)\)\s*;", text):
            decls.append((m.group(1), [a.strip() for a in m.group(2).replace("\n", " ").split(",")]))
    assert len(decls) >= 48
    lines = ["import ctypes as C, sys", "sys.path.insert(0, %r)" % ROOT, "from orb_slam3_fast_b200 import lib", "L = lib.lib()"]
    for name, args in decls:
        vals = []
        for a in args:
            if a in ("void", ""):
                continue
            if "*" in a:
                vals.append("None")
            elif re.search(r"\bfloat\b", a):
                vals.append("C.c_float(0)")
            elif re.search(r"\bdouble\b", a):
                vals.append("C.c_double(0)")
            elif "int64_t" in a:
                vals.append("C.c_int64(0)")
            else:
                vals.append("0")
        lines.append("print(%r, L.%s(%s), flush=True)" % (name, name, ", ".join(vals)))
    r = subprocess.run([sys.executable, "-c", "\n".join(lines)], capture_output=True, text=True)
    answered = dict(l.split() for l in r.stdout.splitlines() if l.startswith("orb"))
    assert r.returncode == 0, "crashed after %s: %s" % (list(answered)[-1:] or "nothing", r.stderr[-500:])
    assert len(answered) == len(decls)
    from orb_slam3_fast_b200 import lib                      # the entry points that return nothing / a string: NULL is a no-op
    L = lib.lib()
    L.orbm_last_error.restype = C.c_char_p
    assert L.orbm_last_error(None) is not None
    L.orbx_host_free(None)
    L.orbm_destroy(None)
    L.orbx_extractor_destroy(None)
    for name, rc in answered.items():
        if name == "orbx_kernel_launches":       # a count, 0 for "no handle"
            assert int(rc) == 0
        else:
            assert int(rc) < 0, (name, rc)


def test_handle_less_entry_points_report_the_missing_device():
    """The front-end entry points take a device ordinal instead of a handle: with valid host buffers and no CUDA device
    they must answer ORBX_E_CUDA (-4) — no CPU fallback, no crash — and pinned allocation must answer NULL."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from orb_slam3_fast_b200 import lib
    L = lib.lib()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    src, dst = np.zeros((48, 64, 3), np.uint8), np.zeros((48, 64), np.uint8)
    mx, my = np.zeros((48, 64), np.float32), np.zeros((48, 64), np.float32)
    assert L.orbx_cvt_gray(0, p(src), 64, 48, 64 * 3, 3, 1, p(dst), 64) == -4
    assert L.orbx_remap_linear(0, p(dst), 64, 48, 64, p(mx), p(my), 64, 48, p(dst), 64) == -4
    assert L.orbx_cvt_gray_device(0, 1, p(src), 64, 48, 64 * 3, C.c_int64(64 * 48 * 3), 3, 1, p(dst), 64,
                                  C.c_int64(64 * 48), None) == -4
    L.orbx_host_alloc.restype = C.c_void_p
    assert L.orbx_host_alloc(C.c_int64(4096)) is None
    handles = (C.c_void_p * 2)(None, None)          # a device list with empty slots is an argument error, not a crash
    assert L.orbx_extract_batch_multi(2, handles, 4, p(dst), 64, 48, 64, C.c_int64(64 * 48), 0, 0, None, None, 0, None,
                                      None) == -3


def test_python_bindings_match_the_headers():
    """The ctypes harness (orb_slam3_fast_b200/lib.py, matcher.py) declares argtypes by hand; most of those calls only run
    on a GPU. Every declared signature must have the header's parameter count, pointers where the header has pointers,
    c_float for float and a 64-bit integer for int64_t — a drifted binding would corrupt arguments silently."""
    import re
    from orb_slam3_fast_b200 import lib
    import orb_slam3_fast_b200.matcher as matcher
    L = lib.lib()
    matcher._bind(L)
    checked = 0
    for h in ("orbx.h", "orbm.h"):
        text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", h)).read(), flags=re.S)
        for m in re.finditer(r"\b(?:int|void\*?|const char\*)\s*\*?\s*(orb[xm]_\w+)\s*\(([^;]*?)\)\s*;", text):
            name = m.group(1)
            params = [a.strip() for a in m.group(2).replace("\n", " ").split(",") if a.strip() not in ("void", "")]
            argtypes = getattr(L, name).argtypes
            if argtypes is None:
                assert name == "orbx_extract_batch_multi", name      # bound at its only call site (extractor.py)
                continue
            assert len(argtypes) == len(params), (name, len(argtypes), len(params))
            for i, (p, t) in enumerate(zip(params, argtypes)):
                is_ptr = "*" in p
                t_ptr = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents")
                assert is_ptr == t_ptr, (name, i, p, t)
                if not is_ptr and re.search(r"\bfloat\b", p):
                    assert t is C.c_float, (name, i, p, t)
                if not is_ptr and "int64_t" in p:
                    assert C.sizeof(t) == 8, (name, i, p, t)
            checked += 1
    assert checked >= 53


def test_python_structures_have_the_layout_of_the_c_header(tmp_path):
    """include/orbx_types.h is mirrored by hand as ctypes Structures twice (orb_slam3_fast_b200/views.py for the product
    harness, oracle/orbref.py for the oracle). A C program generated from the header prints sizeof and every field offset;
    each mirror must have the same size and start every one of its fields on a field boundary of the C struct."""
    import re
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "orbx_types.h")).read(), flags=re.S)
    structs = {}
    for m in re.finditer(r"typedef struct (\w+) \{(.*?)\} \1;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1]
            fields += [re.sub(r"\[\d+\]|[\* ]", "", n) for n in [first] + [n.strip() for n in names[1:]]]
        structs[m.group(1)] = fields
    assert len(structs) >= 13 and structs["orbx_frustum"][:3] == ["Rcw", "tcw", "Ow"]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "orbx_types.h"', "int main(void) {"]
    for name, fields in structs.items():
        src.append('printf("%s %%zu", sizeof(%s));' % (name, name))
        src += ['printf(" %%zu", offsetof(%s, %s));' % (name, f) for f in fields]
        src.append('printf("\\n");')
    src += ["return 0; }"]
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), "-o", exe, str(c)])
    layout = {}
    for line in subprocess.check_output([exe], text=True).splitlines():
        parts = line.split()
        layout[parts[0]] = (int(parts[1]), [int(x) for x in parts[2:]])
    mirror = {"orbx_grid": "Grid", "orbx_frame_view": "FrameView", "orbx_mappoints": "MapPoints", "orbx_fisheye_view": "FisheyeView",
              "orbx_mappoints_right": "MapPointsRight", "orbx_frustum": "Frustum", "orbx_local_map": "LocalMap",
              "orbx_track_params": "TrackParams", "orbx_projected": "Projected", "orbx_featvec": "FeatVec",
              "orbx_vocabulary": "Vocabulary", "orbx_keyframe_view": "KeyFrameView"}
    from orb_slam3_fast_b200 import views
    from oracle import orbref
    compared = 0
    for module in (views, orbref):
        for cname, pyname in mirror.items():
            cls = getattr(module, pyname, None)
            if cls is None:
                continue
            size, offsets = layout[cname]
            assert C.sizeof(cls) == size, (module.__name__, pyname, C.sizeof(cls), size)
            for fname, _ in cls._fields_:
                assert getattr(cls, fname).offset in offsets, (module.__name__, pyname, fname)
            compared += 1
    assert compared >= 22
    from orb_slam3_fast_b200 import synth
    assert synth.frustum(640, 480, seed=0).dtype.itemsize == layout["orbx_frustum"][0]
    assert synth.KP_DTYPE.itemsize == layout["orbx_kp"][0] == 28
