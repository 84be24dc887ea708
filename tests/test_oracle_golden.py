"""Pins the CPU oracle (oracle/cds_oracle.c) on every known-answer vector of the reference's JUnit tests."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import golden_vectors as GV


@pytest.mark.parametrize("case", GV.PIXEL_MATCH, ids=lambda c: f"{c[0]}-{c[1]}")
def test_pixel_match_golden(fixtures, case):
    mask, target, mthr, dthr, ztol, xys, mirror, csw, exp_score, exp_mir = case
    img = fixtures[mask]
    H, W = img.shape[:2]
    m = O.PixelMatchMask(img, mthr, mirror, dthr, ztol, xys, O.label_rects(W, H, csw))
    score, ratio, mirrored = m.score(fixtures[target])
    assert score == exp_score
    assert mirrored == exp_mir
    assert ratio == score / m.size


def test_mask_sizes(fixtures):
    # mask sizes implied by the pixel match vectors
    sizes = {}
    for k in ("em_LPLC2", "em_12191", "em_12191_FL"):
        H, W = fixtures[k].shape[:2]
        sizes[k] = O.PixelMatchMask(fixtures[k], 20, True, 20, 0.01, 2, O.label_rects(W, H)).size
    assert sizes == {"em_LPLC2": 1897, "em_12191": 10299, "em_12191_FL": 17340}


@pytest.mark.parametrize("case", GV.SHAPE_MASK_SIZES)
def test_shape_mask_sizes(fixtures, case):
    mask, thr, exp_qm, exp_he = case
    img = fixtures[mask]
    H, W = img.shape[:2]
    sm = O.ShapeMask(img, thr, True, O.label_rects(W, H))
    assert int(sm.qm.sum()) == exp_qm
    assert int(sm.he.sum()) == exp_he


@pytest.fixture(scope="module")
def shape_masks(fixtures):
    out = {}
    for k in ("em_12191", "em_12191_FL"):
        H, W = fixtures[k].shape[:2]
        out[k] = O.ShapeMask(fixtures[k], 20, True, O.label_rects(W, H))
    return out


@pytest.mark.parametrize("case", GV.SHAPE, ids=lambda c: f"{c[0]}-{c[1]}-{c[2]}-{c[3]}")
def test_shape_golden(fixtures, shape_masks, case):
    mask, target, grad, zgap, exp_gap, exp_he, exp_score, exp_mir = case
    sm = shape_masks[mask]
    t = fixtures[target]
    H, W = t.shape[:2]
    z = fixtures[zgap] if zgap else O.make_zgap(t, 20, O.label_rects(W, H))
    gap, he, mirrored = sm.score(t, fixtures[grad], z)
    assert (gap, he, mirrored) == (exp_gap, exp_he, exp_mir)
    assert O.shape_score_2d(gap, he) == exp_score


def test_shape_missing_variants(fixtures, shape_masks):
    sm = shape_masks["em_12191"]
    assert sm.score(fixtures["lm_BJD"], None, fixtures["zgap_BJD"]) == (-1, -1, False)
    assert sm.score(fixtures["lm_BJD"], fixtures["grad_BJD"], None) == (-1, -1, False)
    assert O.shape_score_2d(-1, -1) == -1


@pytest.mark.parametrize("case", GV.NORMALIZE)
def test_normalize_golden(case):
    pix, gap, he, max_pix, max_neg, exp_shape, exp_norm = case
    shape = O.shape_score_2d(gap, he)
    assert shape == exp_shape
    assert abs(O.normalized_score(pix, shape, max_pix, max_neg) - exp_norm) < 0.1


def test_max_filter_matches_bruteforce():
    rng = np.random.default_rng(7)
    img = np.zeros((41, 67, 3), np.uint8)
    idx = rng.integers(0, 41 * 67, 60)
    img.reshape(-1, 3)[idx] = rng.integers(1, 256, (60, 3))
    for r in (1, 1.5, 2, 2.5, 3, 10, 20):
        a = O.max_filter(img, r)
        b = O.max_filter(img, r, bruteforce=True)
        assert np.array_equal(a, b), r


def test_line_radii_r10():
    # the example in the reference's own doc comment, ImageTransformation.java:541-546
    k, dx = O.line_radii(10)
    assert k == 10
    assert dx.tolist() == [1, 4, 6, 7, 8, 8, 9, 9, 9, 10, 10, 10, 9, 9, 9, 8, 8, 7, 6, 4, 1]


def _all_colours():
    v = np.arange(1 << 24, dtype=np.uint32)
    return np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8)


def test_gray_integer_form():
    # gray == floor((2(r+g+b)+3)/6) for EVERY colour (all 2^24): the integer form the device uses
    # (csrc/cds_shape.cu gray_of) against the oracle's double expression (ColorTransformation.java:40-54)
    cols = _all_colours()
    s = cols.astype(np.int64).sum(axis=1)
    exp = np.where(s == 0, 0, (2 * s + 3) // 6).astype(np.uint8)
    assert np.array_equal(O.rgb_to_gray_batch(cols), exp)


def test_slice_numbers_of_the_lut():
    # GradientAreaGapUtils.java:131-197: a LUT colour is found by exact ratio equality -> its own 1-based index,
    # unless an earlier entry of the same sub-range has the same ratio (none has); black -> 0
    lut = O.lut()
    got = O.slice_numbers(lut.astype(np.uint8))
    assert got.tolist() == list(range(1, 257))
    assert O.slice_number(0, 0, 0) == 0
    # the batch form is the scalar form
    rng = np.random.default_rng(11)
    cols = rng.integers(0, 256, (5000, 3)).astype(np.uint8)
    assert O.slice_numbers(cols).tolist() == [O.slice_number(*c) for c in cols]


def test_is_match():
    assert O.is_match(5, 0.02, 1.0)
    assert not O.is_match(0, 0.5, 1.0)
    assert not O.is_match(5, 0.01, 1.0)   # (float)0.01 > 0.01 is false: float(0.01) < 0.01
    assert O.is_match(1, 1e-9, 0.0)


def test_shift_offsets():
    assert O.shift_offsets(0) == [(0, 0)]
    assert O.shift_offsets(2) == [(-2, -2), (-2, 0), (-2, 2), (0, -2), (0, 0), (0, 2), (2, -2), (2, 0), (2, 2)]
    assert len(O.shift_offsets(4)) == 18 and len(set(O.shift_offsets(4))) == 17
    assert O.lib().cdso_reference_throws_for_xyshift(4) == 1
    assert O.lib().cdso_reference_throws_for_xyshift(2) == 0
    assert O.lib().cdso_reference_throws_for_xyshift(3) == 1
