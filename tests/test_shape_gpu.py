"""GPU parity tests of the shape (gradient area gap) path, through the C ABI, against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O
from tests import golden_vectors as GV

pytestmark = pytest.mark.gpu

W, H = 1210, 566


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


def test_slice_numbers_all_colours(ctx):
    """The device's slice table (csrc/cds_shape.cu slice_of) against GradientAreaGapUtils.calculateSliceGap's slice numbers
    (oracle cdso_slice_number, API/cds/GradientAreaGapUtils.java:18-197) for ALL 2^24 colours."""
    v = np.arange(1 << 24, dtype=np.uint32)
    cols = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8)
    got = ctx.debug_slice_numbers(cols)
    want = O.slice_numbers(cols)
    assert np.array_equal(got, want)
    assert got.max() == 256 and got[0] == 0


def test_mask_sizes_golden(ctx, fixtures):
    # Shape2DMatchColorDepthSearchAlgorithmTest.java:53-54
    for mask, thr, exp_qm, exp_he in GV.SHAPE_MASK_SIZES:
        sms = capi.ShapeMaskSet(ctx, W, H, thr, True, O.label_rects(W, H))
        qm, he = sms.add_rgb(fixtures[mask])
        assert (int(qm[0]), int(he[0])) == (exp_qm, exp_he)
        sms.close()


def test_shape_golden_vectors(ctx, fixtures):
    rects = O.label_rects(W, H)
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    mask_keys = ["em_12191", "em_12191_FL"]
    sms.add_rgb(np.stack([fixtures[k] for k in mask_keys]))
    # the cases with a derived zgap image (zgap_rgb = NULL): targets are (target, gradient) combinations
    derived = [c for c in GV.SHAPE if c[3] is None]
    tg = np.stack([fixtures[c[1]] for c in derived])
    gr = np.stack([fixtures[c[2]] for c in derived])
    pm = [mask_keys.index(c[0]) for c in derived]
    pt = list(range(len(derived)))
    gap, he, mir = sms.score_pairs(tg, gr, None, pm, pt)
    for i, c in enumerate(derived):
        assert (int(gap[i]), int(he[i]), bool(mir[i])) == (c[4], c[5], c[7]), c
        assert capi.shape_score_2d(gap[i], he[i]) == c[6]
    # the case with the on-disk zgap file
    given = [c for c in GV.SHAPE if c[3] is not None]
    tg = np.stack([fixtures[c[1]] for c in given])
    gr = np.stack([fixtures[c[2]] for c in given])
    zg = np.stack([fixtures[c[3]] for c in given])
    gap, he, mir = sms.score_pairs(tg, gr, zg, [mask_keys.index(c[0]) for c in given], list(range(len(given))))
    for i, c in enumerate(given):
        assert (int(gap[i]), int(he[i]), bool(mir[i])) == (c[4], c[5], c[7]), c
    sms.close()


def test_missing_variants_and_errors(ctx, fixtures):
    rects = O.label_rects(W, H)
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    sms.add_rgb(fixtures["em_12191"])
    tg = np.stack([fixtures["lm_BJD"], fixtures["lm_VT033614"]])
    gr = np.stack([fixtures["grad_BJD"], fixtures["grad_VT033614"]])
    gap, he, mir = sms.score_pairs(tg, gr, None, [0, 0], [0, 1], has_variants=[0, 1])
    assert (int(gap[0]), int(he[0]), bool(mir[0])) == (-1, -1, False)       # Shape2DMatch...:155-158
    assert (int(gap[1]), int(he[1])) == (21365, 731)
    assert capi.shape_score_2d(gap[0], he[0]) == -1
    with pytest.raises(capi.CdsIllegalArgument):
        sms.score_pairs(tg, gr, None, [1], [0])                               # mask index out of range
    with pytest.raises(capi.CdsError) as e:
        capi.ShapeMaskSet(ctx, W, H, 20, True, rects, border=3)
    assert e.value.status == capi.CDS_ERR_UNSUPPORTED
    sms.close()


def test_zgap_maker_matches_oracle(ctx, fixtures):
    rects = O.label_rects(W, H)
    imgs = np.stack([fixtures["lm_BJD"], fixtures["lm_GMR"]])
    got = ctx.make_zgap(imgs, 20, 10, rects)
    for i in range(2):
        assert np.array_equal(got[i], O.make_zgap(imgs[i], 20, rects))
    # other radii of the max filter against the oracle's brute force, on a small image
    rng = np.random.default_rng(5)
    small = np.zeros((1, 90, 130, 3), np.uint8)
    idx = rng.integers(0, 90 * 130, 200)
    small.reshape(-1, 3)[idx] = rng.integers(1, 256, (200, 3))
    for r in (1, 2, 2.5, 7, 20, 60):
        got = ctx.make_zgap(small, -1, r, np.zeros((0, 4), np.int32))
        assert np.array_equal(got[0], O.max_filter(small[0], r, bruteforce=(r <= 20))), r


def test_synthetic_pairs_match_oracle(ctx):
    rects = O.label_rects(W, H)
    masks = capi.synth_rgb_host(0, 31, 0, 4, W, H)
    targets = capi.synth_rgb_host(1, 31, 0, 6, W, H)         # target 3 embeds mask 0
    grads = capi.synth_gradient_host(31, 0, 6, W, H)
    for mirror in (True, False):
        sms = capi.ShapeMaskSet(ctx, W, H, 20, mirror, rects)
        qm, he = sms.add_rgb(masks)
        oms = [O.ShapeMask(m, 20, mirror, rects) for m in masks]
        assert qm.tolist() == [int(o.qm.sum()) for o in oms]
        assert he.tolist() == [int(o.he.sum()) for o in oms]
        pm = [m for m in range(4) for _ in range(6)]
        pt = [t for _ in range(4) for t in range(6)]
        gap, hexp, mir = sms.score_pairs(targets, grads, None, pm, pt)
        for i, (m, t) in enumerate(zip(pm, pt)):
            z = O.make_zgap(targets[t], 20, rects)
            assert (int(gap[i]), int(hexp[i]), bool(mir[i])) == oms[m].score(targets[t], grads[t], z), (mirror, m, t)
        # the same pairs with the targets as TIFF files (PackBits and stored, several strip layouts), decoded on the device
        files = [capi.tiff_encode_rgb(t, (8, 566, 1)[i % 3], 32773 if i % 2 else 1) for i, t in enumerate(targets)]
        gap2, hexp2, mir2 = sms.score_pairs_tiff(files, grads, None, pm, pt)
        assert np.array_equal(gap2, gap) and np.array_equal(hexp2, hexp) and np.array_equal(mir2, mir)
        sms.close()


def test_shape_pairs_over_many_tiff_files(ctx):
    """More targets than one upload chunk (32), a has_variants mask, and a bad file that must be named in the error."""
    rects = O.label_rects(W, H)
    masks = capi.synth_rgb_host(0, 77, 0, 3, W, H)
    targets = capi.synth_rgb_host(1, 77, 0, 40, W, H)
    grads = capi.synth_gradient_host(77, 0, 40, W, H)
    files = [capi.tiff_encode_rgb(t, 8, 32773) for t in targets]
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    sms.add_rgb(masks)
    rng = np.random.default_rng(5)
    pm = rng.integers(0, 3, 200)
    pt = rng.integers(0, 40, 200)
    has = (np.arange(40) % 7 != 3).astype(np.uint8)
    a = sms.score_pairs(targets, grads, None, pm, pt, has)
    b = sms.score_pairs_tiff(files, grads, None, pm, pt, has)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    bad = list(files)
    bad[33] = b"II*\0" + b"\0" * 60
    with pytest.raises(capi.CdsError) as e:
        sms.score_pairs_tiff(bad, grads, None, pm, pt, has)
    assert "file 33" in str(e.value)
    c = sms.score_pairs_tiff(files, grads, None, pm, pt, has)            # the context still works
    for x, y in zip(a, c):
        assert np.array_equal(x, y)
    sms.close()


def test_roi_mask_matches_oracle(ctx, fixtures):
    rects = O.label_rects(W, H)
    roi = np.zeros((H, W, 3), np.uint8)
    roi[120:480, 200:760] = 255          # asymmetric on purpose: the ROI is not mirrored with the query
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects, roi=roi)
    sms.add_rgb(fixtures["em_12191"])
    om = O.ShapeMask(fixtures["em_12191"], 20, True, rects, roi=roi)
    for tkey, gkey in (("lm_VT016795", "grad_VT016795"), ("lm_BJD", "grad_BJD")):
        t = fixtures[tkey]
        z = O.make_zgap(t, 20, rects)
        gap, he, mir = sms.score_pairs(t[None], fixtures[gkey][None], z[None], [0], [0])
        assert (int(gap[0]), int(he[0]), bool(mir[0])) == om.score(t, fixtures[gkey], z)
    sms.close()


@pytest.mark.parametrize("size", [(333, 151), (96, 64), (1025, 70)], ids=["odd", "one_strip", "wide"])
def test_small_and_odd_image_sizes_match_oracle(ctx, size):
    """Image sizes that are not the library's: odd widths (unaligned slice-plane rows), a single strip, partly filled last tiles;
    random sparse colours next to the borders and inside a label rectangle, derived and given zgap images, with and without ROI."""
    Ws, Hs = size
    rng = np.random.default_rng(Ws * 1000 + Hs)
    rects = np.array([[0, 0, min(40, Ws // 3), 12], [Ws - 17, Hs - 9, Ws, Hs]], np.int32)

    def sparse(n, lo=1):
        img = np.zeros((Hs, Ws, 3), np.uint8)
        idx = rng.integers(0, Hs * Ws, n)
        img.reshape(-1, 3)[idx] = rng.integers(lo, 256, (n, 3))
        img[0, :5] = 200; img[-1, -5:] = 90; img[:4, -1] = 33          # pixels on the borders
        return img

    dim = np.zeros((Hs, Ws, 3), np.uint8)               # channel values 0..2: the gray(max60) > 0 test is "channel maxima add up to >= 2"
    idx = rng.integers(0, Hs * Ws, 40)
    dim.reshape(-1, 3)[idx] = rng.integers(0, 3, (40, 3))
    masks = np.stack([sparse(60 + 40 * i, 0) for i in range(3)] + [dim])
    targets = np.stack([sparse(150 + 100 * i) for i in range(5)])
    grads = rng.integers(0, 700, (5, Hs, Ws)).astype(np.uint16)
    roi = np.zeros((Hs, Ws, 3), np.uint8)
    roi[Hs // 5: Hs - 3, Ws // 7: Ws - Ws // 9] = 255
    for r in (None, roi):
        sms = capi.ShapeMaskSet(ctx, Ws, Hs, 20, True, rects, roi=r)
        qm, he = sms.add_rgb(masks)
        oms = [O.ShapeMask(m, 20, True, rects, roi=r) for m in masks]
        assert qm.tolist() == [int(o.qm.sum()) for o in oms]
        assert he.tolist() == [int(o.he.sum()) for o in oms]
        pm = [m for m in range(4) for _ in range(5)]
        pt = [t for _ in range(4) for t in range(5)]
        zg = np.stack([O.make_zgap(t, 20, rects) for t in targets])
        got_z = ctx.make_zgap(targets, 20, 10, rects)
        assert np.array_equal(got_z, zg)
        for z in (None, zg):
            gap, hexp, mir = sms.score_pairs(targets, grads, z, pm, pt)
            for i, (m, t) in enumerate(zip(pm, pt)):
                assert (int(gap[i]), int(hexp[i]), bool(mir[i])) == oms[m].score(targets[t], grads[t], zg[t]), (size, r is not None, z is not None, m, t)
        sms.close()
