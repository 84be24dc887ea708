"""ctypes binding of the CPU oracle (oracle/cds_oracle.c).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py, never from colormipsearch_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcdsoracle.so")

# label regions of the reference tests / CLI for a W-wide image
# (colormipsearch-api/src/test/java/org/janelia/colormipsearch/ImageTestUtils.java:12-24,
#  colormipsearch-tools/src/main/java/org/janelia/colormipsearch/cmd/AbstractColorDepthMatchArgs.java:101-119)


def label_rects(W, H, color_scale_width=270):
    rects = []
    if W > color_scale_width:
        rects.append((W - color_scale_width, 0, W, 90))
    rects.append((0, 0, 330, 100))
    return np.asarray(rects, dtype=np.int32).reshape(-1, 4)


def build(force=False):
    src = os.path.join(_HERE, "cds_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libcdsoracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p = C.POINTER(C.c_uint8)
        u16p = C.POINTER(C.c_uint16)
        i32p = C.POINTER(C.c_int32)
        i64p = C.POINTER(C.c_int64)
        L.cdso_pixel_gap.restype = C.c_double
        L.cdso_pixel_gap.argtypes = [C.c_int] * 6
        L.cdso_mask_positions.restype = C.c_int
        L.cdso_mask_positions.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, C.c_int, i32p, i32p]
        L.cdso_shift_offsets.restype = C.c_int
        L.cdso_shift_offsets.argtypes = [C.c_int, i32p, i32p, C.c_int]
        L.cdso_reference_throws_for_xyshift.restype = C.c_int
        L.cdso_reference_throws_for_xyshift.argtypes = [C.c_int]
        L.cdso_mask_create.restype = C.c_void_p
        L.cdso_mask_create.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                       i32p, C.c_int]
        L.cdso_mask_destroy.argtypes = [C.c_void_p]
        L.cdso_mask_size.restype = C.c_int
        L.cdso_mask_size.argtypes = [C.c_void_p]
        L.cdso_mask_nvariants.restype = C.c_int
        L.cdso_mask_nvariants.argtypes = [C.c_void_p]
        L.cdso_mask_get_positions.argtypes = [C.c_void_p, i32p]
        L.cdso_mask_get_bbox.argtypes = [C.c_void_p, i32p]
        L.cdso_mask_score.restype = C.c_int
        L.cdso_mask_score.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, i32p, C.POINTER(C.c_double), i32p]
        L.cdso_mask_variant_scores.argtypes = [C.c_void_p, u8p, i32p]
        L.cdso_is_match.restype = C.c_int
        L.cdso_is_match.argtypes = [C.c_int, C.c_double, C.c_double]
        L.cdso_search_dense.restype = C.c_int
        L.cdso_search_dense.argtypes = [C.POINTER(C.c_void_p), C.c_int, u8p, C.c_int64, C.c_int, i32p, u8p]
        L.cdso_lut.argtypes = [i32p]
        L.cdso_slice_number.restype = C.c_int
        L.cdso_slice_number.argtypes = [C.c_int] * 3
        L.cdso_slice_numbers.argtypes = [u8p, C.c_int64, u16p]
        L.cdso_rgb_to_gray_batch.argtypes = [u8p, C.c_int64, u8p]
        L.cdso_slice_gap.restype = C.c_int
        L.cdso_slice_gap.argtypes = [C.c_int, C.c_int]
        L.cdso_shape_score_2d.restype = C.c_int64
        L.cdso_shape_score_2d.argtypes = [C.c_int64, C.c_int64]
        L.cdso_normalized_score.restype = C.c_double
        L.cdso_normalized_score.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_int64]
        L.cdso_rgb_to_gray.restype = C.c_int
        L.cdso_rgb_to_gray.argtypes = [C.c_int] * 3
        L.cdso_mask_rgb.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
        L.cdso_clear_regions.argtypes = [u8p, C.c_int, C.c_int, i32p, C.c_int, u8p]
        L.cdso_line_radii.restype = C.c_int
        L.cdso_line_radii.argtypes = [C.c_double, i32p, C.c_int]
        L.cdso_max_filter_bruteforce.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, u8p]
        L.cdso_max_filter.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_double, u8p]
        L.cdso_shape_prepare_mask.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, C.c_int, u8p, u8p, u8p]
        L.cdso_shape_score.restype = C.c_int
        L.cdso_shape_score.argtypes = [u8p, u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       i32p, C.c_int, u8p, u16p, u8p, i64p, i64p, i32p]
        L.cdso_make_zgap.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, C.c_int, u8p]
        L.cdso_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _u16(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _rects(rects):
    r = np.ascontiguousarray(np.asarray(rects, dtype=np.int32).reshape(-1, 4))
    return r, _i32(r), int(r.shape[0])


def _rgb(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3, a.shape
    return a


def pixel_gap(c1, c2):
    return lib().cdso_pixel_gap(int(c1[0]), int(c1[1]), int(c1[2]), int(c2[0]), int(c2[1]), int(c2[2]))


def shift_offsets(xyshift):
    cap = 9 * (xyshift // 2) + 1
    dx = np.zeros(cap, np.int32)
    dy = np.zeros(cap, np.int32)
    n = lib().cdso_shift_offsets(xyshift, _i32(dx), _i32(dy), cap)
    return list(zip(dx[:n].tolist(), dy[:n].tolist()))


class PixelMatchMask:
    """A prepared query = PixelMatchColorDepthSearchAlgorithm instance
    (colormipsearch-api/src/main/java/org/janelia/colormipsearch/cds/PixelMatchColorDepthSearchAlgorithm.java:29-101)."""

    def __init__(self, rgb, query_threshold, mirror, target_threshold, z_tolerance, xyshift, rects):
        rgb = _rgb(rgb)
        self.H, self.W = rgb.shape[:2]
        r, rp, nr = _rects(rects)
        self._h = lib().cdso_mask_create(_u8(rgb), self.W, self.H, int(query_threshold), int(bool(mirror)),
                                         int(target_threshold), float(z_tolerance), int(xyshift), rp, nr)
        self.size = lib().cdso_mask_size(self._h)
        self.nvariants = lib().cdso_mask_nvariants(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cdso_mask_destroy(self._h)
            self._h = None

    def positions(self):
        out = np.zeros(max(self.size, 1), np.int32)
        lib().cdso_mask_get_positions(self._h, _i32(out))
        return out[: self.size]

    def bbox(self):
        out = np.zeros(4, np.int32)
        lib().cdso_mask_get_bbox(self._h, _i32(out))
        return out

    def score(self, target):
        target = _rgb(target)
        s = C.c_int32()
        m = C.c_int32()
        r = C.c_double()
        rc = lib().cdso_mask_score(self._h, _u8(target), target.shape[1], target.shape[0],
                                   C.byref(s), C.byref(r), C.byref(m))
        if rc != 0:
            raise ValueError("Invalid image size - target's image size must match query's image size")
        return s.value, r.value, bool(m.value)

    def variant_scores(self, target):
        target = _rgb(target)
        out = np.zeros(self.nvariants, np.int32)
        lib().cdso_mask_variant_scores(self._h, _u8(target), _i32(out))
        return out


def search_dense(masks, targets, nthreads=0):
    """masks: list[PixelMatchMask]; targets: uint8 [T,H,W,3] -> (scores int32 [M,T], mirrored uint8 [M,T], nthreads)"""
    targets = np.ascontiguousarray(targets, dtype=np.uint8)
    T = targets.shape[0]
    M = len(masks)
    scores = np.zeros((M, T), np.int32)
    mirrored = np.zeros((M, T), np.uint8)
    handles = (C.c_void_p * max(M, 1))(*[m._h for m in masks])
    used = lib().cdso_search_dense(handles, M, _u8(targets), T, int(nthreads), _i32(scores), _u8(mirrored))
    return scores, mirrored, used


def is_match(score, ratio, pct_positive_pixels):
    return bool(lib().cdso_is_match(int(score), float(ratio), float(pct_positive_pixels)))


def lut():
    out = np.zeros((256, 3), np.int32)
    lib().cdso_lut(_i32(out))
    return out


def slice_number(r, g, b):
    return lib().cdso_slice_number(int(r), int(g), int(b))


def slice_numbers(rgb):
    """slice number of every colour of rgb[n][3] (uint16[n])"""
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    out = np.empty(len(rgb), np.uint16)
    lib().cdso_slice_numbers(_u8(rgb), len(rgb), _u16(out))
    return out


def rgb_to_gray_batch(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8).reshape(-1, 3)
    out = np.empty(len(rgb), np.uint8)
    lib().cdso_rgb_to_gray_batch(_u8(rgb), len(rgb), _u8(out))
    return out


def slice_gap(rgb1, rgb2):
    return lib().cdso_slice_gap(int(rgb1), int(rgb2))


def shape_score_2d(gap, he):
    return lib().cdso_shape_score_2d(int(gap), int(he))


def normalized_score(pix, shape, max_pix, max_shape):
    return lib().cdso_normalized_score(int(pix), int(shape), int(max_pix), int(max_shape))


def rgb_to_gray(r, g, b):
    return lib().cdso_rgb_to_gray(int(r), int(g), int(b))


def mask_rgb(img, threshold):
    img = _rgb(img)
    out = np.empty_like(img)
    lib().cdso_mask_rgb(_u8(img), img.shape[1], img.shape[0], int(threshold), _u8(out))
    return out


def clear_regions(img, rects):
    img = _rgb(img)
    out = np.empty_like(img)
    r, rp, nr = _rects(rects)
    lib().cdso_clear_regions(_u8(img), img.shape[1], img.shape[0], rp, nr, _u8(out))
    return out


def line_radii(radius):
    k = lib().cdso_line_radii(float(radius), None, 0)
    dx = np.zeros(2 * k + 1, np.int32)
    lib().cdso_line_radii(float(radius), _i32(dx), 2 * k + 1)
    return k, dx


def max_filter(img, radius, bruteforce=False):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    nch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    fn = lib().cdso_max_filter_bruteforce if bruteforce else lib().cdso_max_filter
    fn(_u8(img), img.shape[1], img.shape[0], nch, float(radius), _u8(out))
    return out


def make_zgap(target, threshold, rects):
    target = _rgb(target)
    out = np.empty_like(target)
    r, rp, nr = _rects(rects)
    lib().cdso_make_zgap(_u8(target), target.shape[1], target.shape[0], int(threshold), rp, nr, _u8(out))
    return out


class ShapeMask:
    """A prepared query of the shape score = what createShapeMatchCDSAlgorithmProvider builds per mask
    (colormipsearch-api/src/main/java/org/janelia/colormipsearch/cds/ColorDepthSearchAlgorithmProviderFactory.java:76-127)."""

    def __init__(self, rgb, query_threshold, mirror, rects, border=0, roi=None):
        rgb = _rgb(rgb)
        self.H, self.W = rgb.shape[:2]
        self.query_threshold = int(query_threshold)
        self.mirror = bool(mirror)
        self.border = int(border)
        self.rects = np.ascontiguousarray(np.asarray(rects, np.int32).reshape(-1, 4))
        self.q = np.empty_like(rgb)
        self.qm = np.empty((self.H, self.W), np.uint8)
        self.he = np.empty((self.H, self.W), np.uint8)
        r, rp, nr = _rects(self.rects)
        lib().cdso_shape_prepare_mask(_u8(rgb), self.W, self.H, self.border, rp, nr,
                                      _u8(self.q), _u8(self.qm), _u8(self.he))
        # ROI image is label-cleared too (:97-101)
        self.roi = None if roi is None else clear_regions(roi, self.rects)

    def score(self, target, grad, zgap):
        """-> (gradientAreaGap, highExpressionArea, mirrored); (-1, -1, False) when a variant is missing."""
        target = _rgb(target)
        gap = C.c_int64()
        he = C.c_int64()
        mir = C.c_int32()
        gp = None
        zp = None
        if grad is not None:
            grad = np.ascontiguousarray(grad).astype(np.uint16, copy=False)
            grad = np.ascontiguousarray(grad)
            gp = _u16(grad)
        if zgap is not None:
            zgap = _rgb(zgap)
            zp = _u8(zgap)
        r, rp, nr = _rects(self.rects)
        lib().cdso_shape_score(_u8(self.q), _u8(self.qm), _u8(self.he),
                               None if self.roi is None else _u8(self.roi),
                               self.W, self.H, self.border, self.query_threshold, int(self.mirror),
                               rp, nr, _u8(target), gp, zp, C.byref(gap), C.byref(he), C.byref(mir))
        return gap.value, he.value, bool(mir.value)


def num_threads():
    return lib().cdso_num_threads()


# ------------------------------------------------------------------------------------------------------------------
# f1: the selection of the matches that go on to shape scoring.  Pure-Python restatement (small inputs) of
#   ItemsHandling.selectTopRankedElements   colormipsearch-api/src/main/java/org/janelia/colormipsearch/results/ItemsHandling.java:80-109
#   ColorMIPProcessUtils.selectBestMatches   colormipsearch-tools/src/main/java/org/janelia/colormipsearch/cmd/cdsprocess/ColorMIPProcessUtils.java:12-34
# including the iteration order of the java.util.HashMap that Collectors.groupingBy fills (ties between groups keep it).
# ------------------------------------------------------------------------------------------------------------------
def java_string_hash(s):
    """java.lang.String.hashCode()."""
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h


class JavaHashMap:
    """Insertion and iteration order of java.util.HashMap<String, list> (OpenJDK 8+): table of 16 buckets, doubled when
    ++size > 0.75 * capacity, resize splits every bucket preserving order; bins are kept as lists (no treeification)."""

    def __init__(self):
        self.table = [[] for _ in range(16)]
        self.size = 0

    @staticmethod
    def _spread(key):
        h = java_string_hash(key)
        return h ^ (h >> 16)

    def get_or_create(self, key):
        b = self.table[self._spread(key) & (len(self.table) - 1)]
        for k, v in b:
            if k == key:
                return v
        v = []
        b.append((key, v))
        self.size += 1
        if self.size > 0.75 * len(self.table):
            old, cap = self.table, 2 * len(self.table)
            self.table = [[] for _ in range(cap)]
            for bucket in old:                      # lo / hi split keeps relative order
                for k, val in bucket:
                    self.table[self._spread(k) & (cap - 1)].append((k, val))
        return v

    def entries(self):
        return [(k, v) for bucket in self.table for k, v in bucket]


def select_top_ranked(items, key_fn, score_fn, top_results, limit_sub_results):
    """ItemsHandling.selectTopRankedElements -> list of (key, best score, sorted items)."""
    groups = JavaHashMap()
    for it in items:
        k = key_fn(it)
        if k is None or not k.strip():
            k = "UNKNOWN"                            # StringUtils.defaultIfBlank
        groups.get_or_create(k).append(it)
    scored = []
    for k, r in groups.entries():
        r.sort(key=lambda it: -float(score_fn(it)))      # List.sort(comparator.reversed()): stable
        if 0 < limit_sub_results < len(r):
            r = r[:limit_sub_results]
        scored.append((k, max(float(score_fn(it)) for it in r), r))
    scored.sort(key=lambda e: -e[1])                     # Stream.sorted: stable on HashMap order
    if top_results > 0 and len(scored) > top_results:
        scored = scored[:top_results]
    return scored


def select_best_matches(matches, top_lines, top_samples_per_line, top_matches_per_sample):
    """matches: list of (line, sample, score, ...).  Returns the kept matches in the reference's output order."""
    out = []
    for _, _, line_items in select_top_ranked(matches, lambda m: m[0], lambda m: m[2], top_lines, -1):
        for _, _, sample_items in select_top_ranked(line_items, lambda m: m[1], lambda m: m[2], top_samples_per_line, top_matches_per_sample):
            out.extend(sample_items)
    return out
