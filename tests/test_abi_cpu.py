"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/cdsgpu.h declares, fails loudly without a
device, and its host-only entry points (score post-processing, match-interval tables, synthetic generator) are right."""
import os
import re
import subprocess

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O
from tests import golden_vectors as GV

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "cdsgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cds_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cds_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    # the Python binding knows exactly the same entry points
    assert sorted(capi.SIGNATURES) == names
    capi.lib()
    assert capi.lib().cds_abi_version() == 2


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(capi.CdsError) as e:
        capi.Context()
    assert e.value.status == capi.CDS_ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "colormipsearch_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "cdso_" not in text and "libcdsoracle" not in text and "from oracle" not in text and "import oracle" not in text, f


@pytest.mark.parametrize("case", GV.NORMALIZE)
def test_score_postprocessing_golden(case):
    pix, gap, he, max_pix, max_neg, exp_shape, exp_norm = case
    shape = capi.shape_score_2d(gap, he)
    assert shape == exp_shape
    got = capi.normalized_score(pix, shape, max_pix, max_neg)
    assert abs(got - exp_norm) < 0.1
    ref = O.normalized_score(pix, shape, max_pix, max_neg)
    assert got == ref or abs(got - ref) <= 1e-6 * abs(ref)


def test_normalize_scores_matches_oracle():
    rng = np.random.default_rng(5)
    pix = rng.integers(0, 900, 200).astype(np.int32)
    gaps = rng.integers(-1, 200000, 200).astype(np.int64)
    hes = rng.integers(-1, 30000, 200).astype(np.int64)
    got = capi.normalize_scores(pix, gaps, hes)
    shapes = [O.shape_score_2d(g, h) for g, h in zip(gaps, hes)]
    max_pix, max_shape = int(pix.max()), max(shapes)
    exp = np.array([np.float32(O.normalized_score(p, s, max_pix, max_shape)) for p, s in zip(pix, shapes)], np.float32)
    assert np.allclose(got, exp, rtol=1e-6, atol=0)
    assert capi.shape_score_2d(-1, 5) == -1 and capi.shape_score_2d(5, -1) == -1
    assert capi.normalized_score(0, 10, 5, 5) == 0 and capi.normalized_score(7, -1, 9, 5) == 7


def _ratio_ranks():
    a, b = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    valid = a < b
    ratios = np.where(valid, a / np.maximum(b, 1), 0.0)
    uniq = np.unique(ratios[valid])
    assert len(uniq) == 19820
    return uniq


_COLOR_OF = {  # sector -> (second, max, third) -> (r, g, b)
    0: lambda a, m, t: (a, t, m), 1: lambda a, m, t: (t, a, m), 2: lambda a, m, t: (t, m, a),
    3: lambda a, m, t: (a, m, t), 4: lambda a, m, t: (m, a, t), 5: lambda a, m, t: (m, t, a),
}


@pytest.mark.parametrize("tol", [0.01, 0.02, 0.005, 0.0, 0.25, -0.001])
def test_match_intervals_equal_double_predicate(tol):
    """The two-interval integer predicate == (calculatePixelGap <= zTolerance) of the oracle, for sampled mask colours
    against one representative colour of EVERY target class."""
    uniq = _ratio_ranks()
    rng = np.random.default_rng(11)
    # representative (second, max) per rank
    rep = {}
    for m in range(255, 0, -1):
        for a in range(m):
            rep.setdefault(a / m, (a, m))
    reps = [rep[v] for v in uniq]
    gap = O.lib().cdso_pixel_gap
    for _ in range(12):
        s1 = int(rng.integers(0, 6))
        k1 = int(rng.integers(0, len(uniq)))
        a1, m1 = reps[k1]
        strict = s1 in (0, 2, 4)
        if strict and a1 == 0:
            continue
        t1 = int(rng.integers(0, a1)) if strict else int(rng.integers(0, a1 + 1))
        c1 = _COLOR_OF[s1](a1, m1, t1)
        lo1, len1, lo2, len2 = capi.class_intervals(tol, s1, k1)
        for s2 in range(6):
            strict2 = s2 in (0, 2, 4)
            for k2 in range(0, len(uniq), 7):
                a2, m2 = reps[k2]
                if strict2 and a2 == 0:
                    continue
                c2 = _COLOR_OF[s2](a2, m2, 0)
                expect = gap(*c1, *c2) <= tol
                sr = s2 * 32768 + k2
                got = (lo1 != 0xFFFFFFFF and 0 <= sr - lo1 <= len1) or (lo2 != 0xFFFFFFFF and 0 <= sr - lo2 <= len2)
                assert expect == got, (tol, c1, c2, (lo1, len1, lo2, len2))


def test_synth_host_generator_properties():
    W, H = 1210, 566
    masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 12, W, H)
    again = capi.synth_rgb_host(0, 0xC0FFEE, 5, 2, W, H)
    assert np.array_equal(masks[5:7], again)                      # pure function of (kind, seed, index)
    rects = O.label_rects(W, H)
    sizes = [O.PixelMatchMask(m, 20, False, 20, 0.01, 0, rects).size for m in masks]
    assert min(sizes) > 300 and max(sizes) < 60000, sizes
    # labels are drawn only inside the excluded regions
    for m in masks:
        cleared = O.clear_regions(m, rects)
        assert (m != cleared).any()
    targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 6, W, H)
    cov = [(t.max(axis=2) > 20).mean() for t in targets]
    assert 0.001 < min(cov) and max(cov) < 0.25, cov
    assert not np.array_equal(targets[0], targets[1])
    # target 3 embeds a jittered copy of mask 0: it must score far better against mask 0 than its neighbours do
    m0 = O.PixelMatchMask(masks[0], 20, True, 20, 0.02, 2, rects)
    s = [m0.score(t)[0] for t in targets]
    assert s[3] > 5 * (max(s[:3] + s[4:]) + 1), s
    g = capi.synth_gradient_host(0xC0FFEE, 3, 1, W, H)[0]
    sig = targets[3].max(axis=2) > 0
    inside = np.ones((H, W), bool)
    for r in rects:
        inside[r[1]:r[3], r[0]:r[2]] = False
    assert g[sig & inside].max() == 0 and g.max() <= 650 and g.max() > 50


def test_host_only_generator_equals_the_library_generator():
    """oracle/libcdssynth.so (the reference arm's input generator, built from the same header without CUDA) renders the images the
    GPU library renders."""
    from oracle import synth as S
    for kind, idx in ((0, 3), (1, 11), (1, 20)):
        assert np.array_equal(S.synth_rgb(kind, 0xC0FFEE, idx, 1, 1210, 566), capi.synth_rgb_host(kind, 0xC0FFEE, idx, 1, 1210, 566))
    assert np.array_equal(S.synth_gradient(0xC0FFEE, 5, 1, 301, 97), capi.synth_gradient_host(0xC0FFEE, 5, 1, 301, 97))


def test_profile_tool_markers_exist_in_the_kernel_source():
    """tools/ncu_phases.py groups an ncu report by marker lines of cds_cand.cu; a marker that an edit removed must fail here, not when the
    next profile is read."""
    import importlib.util
    path = os.path.join(ROOT, "tools", "ncu_phases.py")
    spec = importlib.util.spec_from_file_location("ncu_phases", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)                     # raises SystemExit("marker not found ...") when a marker is gone
    lines = [l for l, _ in mod.OUTER]
    assert lines == sorted(lines) and len(set(lines)) == len(lines)


def test_every_context_option_is_documented_in_the_header():
    """cds_ctx_set_option's names live in csrc/cds_api.cu; include/cdsgpu.h is where a caller learns about them."""
    import re
    src = open(os.path.join(ROOT, "colormipsearch_b200", "csrc", "cds_api.cu")).read()
    body = src[src.index('extern "C" cds_status cds_ctx_set_option'):]
    body = body[:body.index("unknown option")]
    names = set(re.findall(r'std::strcmp\(name, "([a-z_0-9]+)"\)', body))
    assert {"match_kernel", "stream_chunk", "device_inflate", "wide_lists", "shape_inflate_window"} <= names
    header = open(os.path.join(ROOT, "include", "cdsgpu.h")).read()
    doc = header[header.index("Tuning / test switches"):header.index("cds_status cds_ctx_set_option")]
    missing = sorted(n for n in names if '"%s"' % n not in doc)
    assert not missing, missing


def test_ctypes_structs_follow_the_header():
    """The structs that cross the ABI by value-in-memory: same fields, in the same order, with the same widths as include/cdsgpu.h."""
    import ctypes as C
    import re
    header = open(os.path.join(ROOT, "include", "cdsgpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    widths = {"int32_t": 4, "uint32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8, "float": 4, "uint8_t": 1}

    def fields_of(name):
        m = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), header, flags=re.S)
        assert m, name
        out = []
        for decl in m.group(1).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            typ, rest = decl.split(None, 1)
            for f in rest.split(","):
                out.append((f.strip(), widths[typ]))
        return out

    for cname, ctype in (("cds_search_stats", capi.SearchStats), ("cds_tiff_info", capi.TiffInfo), ("cds_png_info", capi.PngInfo),
                         ("cds_zip_entry", capi.ZipEntry)):
        want = fields_of(cname)
        got = [(n, C.sizeof(t)) for n, t in ctype._fields_]
        assert got == want, (cname, got, want)


def test_ctypes_signatures_follow_the_header():
    """Every function the header declares has a Python prototype with the same number of arguments, scalar arguments of the same width
    in the same places, and pointers where the header has pointers."""
    import ctypes as C
    import re
    header = open(os.path.join(ROOT, "include", "cdsgpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    header = re.sub(r"typedef struct \w+ \{.*?\} \w+;", "", header, flags=re.S)
    scalar = {"int32_t": C.c_int32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "double": C.c_double, "float": C.c_float, "int": C.c_int32}
    decls = re.findall(r"^\s*(?:const\s+)?(\w+)\s*(\*?)\s*(cds_\w+)\s*\(([^;{]*?)\)\s*;", header, flags=re.M)
    assert len(decls) > 40
    checked = 0
    for ret, retptr, name, args in decls:
        assert name in capi.SIGNATURES, "no Python prototype for " + name
        restype, argtypes = capi.SIGNATURES[name]
        args = [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == len(argtypes), (name, len(args), len(argtypes))
        for decl, ct in zip(args, argtypes):
            if "*" in decl:
                assert ct is C.c_void_p or ct is C.c_char_p or hasattr(ct, "contents") or issubclass(ct, C._Pointer), (name, decl, ct)
            else:
                typ = [w for w in decl.split() if w != "const"][0]
                if typ in scalar:
                    assert C.sizeof(ct) == C.sizeof(scalar[typ]), (name, decl, ct)
        checked += 1
    assert checked == len(decls)
