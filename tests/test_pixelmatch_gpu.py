"""GPU parity tests of the pixel-match path, through the C ABI, against the CPU oracle and the reference's golden vectors."""
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


def _np_expected_codes(rgb, thr):
    """numpy restatement of the code word (cds_common.h) from the sector ladder of calculatePixelGap."""
    r, g, b = (rgb[:, i].astype(np.int64) for i in range(3))
    a_, b_ = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    valid = a_ < b_
    uniq = np.unique((a_ / np.maximum(b_, 1))[valid])
    rank_tab = np.searchsorted(uniq, a_ / np.maximum(b_, 1))
    sector = np.full(len(r), -1)
    second = np.zeros(len(r), np.int64)
    mx = np.maximum(np.maximum(r, g), b)
    bmax = (b > r) & (b > g)
    gmax = (g > b) & (g > r) & ~bmax
    rmax = (r > b) & (r > g) & ~bmax & ~gmax
    sector[bmax & (r > g)] = 0; second[bmax & (r > g)] = r[bmax & (r > g)]
    sector[bmax & ~(r > g)] = 1; second[bmax & ~(r > g)] = g[bmax & ~(r > g)]
    sector[gmax & (b > r)] = 2; second[gmax & (b > r)] = b[gmax & (b > r)]
    sector[gmax & ~(b > r)] = 3; second[gmax & ~(b > r)] = r[gmax & ~(b > r)]
    sector[rmax & (g > b)] = 4; second[rmax & (g > b)] = g[rmax & (g > b)]
    sector[rmax & ~(g > b)] = 5; second[rmax & ~(g > b)] = b[rmax & ~(g > b)]
    sr = np.where(sector >= 0, sector * 32768 + rank_tab[second, np.maximum(mx, 1)], 6 * 32768)
    code = (sr.astype(np.uint64) << 8) | mx.astype(np.uint64)
    code = np.where(mx > thr, code, code | 0x80000000)
    return code.astype(np.uint32)


def test_encode_all_16M_colours(ctx):
    v = np.arange(1 << 24, dtype=np.uint32)
    rgb = np.stack([(v >> 16) & 255, (v >> 8) & 255, v & 255], axis=1).astype(np.uint8)
    for thr in (20, 100):
        got = ctx.debug_encode_colors(rgb, thr)
        exp = _np_expected_codes(rgb, thr)
        assert np.array_equal(got, exp)


BATCHED_KERNELS = ("cand", "band")     # the two batched kernels; fewer than 16 masks always take the gather kernel


def _maskset(ctx, case_or_params, rects):
    mthr, dthr, ztol, xys, mirror = case_or_params
    return capi.MaskSet(ctx, W, H, mthr, dthr, ztol, xys, mirror, rects)


@pytest.mark.parametrize("case", GV.PIXEL_MATCH, ids=lambda c: f"{c[0]}-{c[1]}")
def test_golden_pair_call(ctx, fixtures, case):
    mask, target, mthr, dthr, ztol, xys, mirror, csw, exp_score, exp_mir = case
    ms = _maskset(ctx, (mthr, dthr, ztol, xys, mirror), O.label_rects(W, H, csw))
    sizes = ms.add_rgb(fixtures[mask])
    score, ratio, mirrored = ms.score_pair(0, fixtures[target])
    assert (score, mirrored) == (exp_score, exp_mir)
    assert ratio == score / sizes[0]
    ms.close()


def test_golden_dense_both_kernels(ctx, fixtures):
    """All five provider vectors in one dense search; the mask set is padded with synthetic masks so that the batched
    band kernel runs, and a 2-mask set exercises the gather kernel on the same library."""
    rects = O.label_rects(W, H, 270)
    lm_keys = ["lm_VT033614", "lm_BJD", "lm_VT016795", "lm_GMR"]
    lib = capi.Library(ctx, W, H, 16)
    lib.add_rgb(np.stack([fixtures[k] for k in lm_keys]))
    extra = capi.synth_rgb_host(0, 77, 0, 18, W, H)
    masks = np.concatenate([np.stack([fixtures["em_12191"], fixtures["em_12191_FL"]]), extra])
    for n_masks, kern in ((20, "cand"), (20, "band"), (2, "auto")):
        ctx.set_match_kernel(kern)
        ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
        sizes = ms.add_rgb(masks[:n_masks])
        assert sizes[0] == 10299 and sizes[1] == 17340
        scores, mirrored = ms.search_dense(lib)
        assert ctx.last_stats()["match_kernel"] == {"cand": 1, "band": 2, "auto": 3}[kern]
        exp = {(0, 0): (439, 0), (0, 1): (414, 0), (1, 0): (515, 0), (1, 2): (483, 0), (0, 2): (426, 1)}
        for (m, t), (s, mir) in exp.items():
            assert (scores[m, t], mirrored[m, t]) == (s, mir), (n_masks, m, t)
        # and every cell against the oracle
        oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in masks[:n_masks]]
        es, em, _ = O.search_dense(oms, np.stack([fixtures[k] for k in lm_keys]))
        assert np.array_equal(scores, es)
        assert np.array_equal(mirrored, em)
        ms.close()
    ctx.set_match_kernel("auto")
    lib.close()


@pytest.fixture(scope="module")
def synth(ctx):
    masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 24, W, H)
    targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 70, W, H)   # 70 > one library block of 64
    lib = capi.Library(ctx, W, H, 80)
    lib.add_rgb(targets)
    yield masks, targets, lib
    lib.close()


PARAMS = [
    # (maskThr, dataThr, zTol, xyShift, mirror)
    (20, 20, 0.01, 2, True),      # production (cdsparams.sh)
    (100, 100, 0.02, 0, True),    # CLI defaults + mirror (BASELINE config 1)
    (100, 100, 0.02, 0, False),
    (20, 20, 0.005, 4, True),     # BASELINE config 4 (the Java reference throws for xyShift 4; the oracle is the spec)
    (20, 40, 0.01, 2, False),
    (20, 20, 0.01, 6, True),      # gather kernel only
]


@pytest.mark.parametrize("params", PARAMS, ids=lambda p: "thr%d-%d_tol%g_xy%d_mir%d" % p)
@pytest.mark.parametrize("n_masks", [24, 3])
def test_dense_matches_oracle_on_synthetic(ctx, synth, params, n_masks):
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    mthr, dthr, ztol, xys, mirror = params
    ms = _maskset(ctx, params, rects)
    sizes = ms.add_rgb(masks[:n_masks])
    oms = [O.PixelMatchMask(x, mthr, mirror, dthr, ztol, xys, rects) for x in masks[:n_masks]]
    assert sizes.tolist() == [m.size for m in oms]
    es, em, _ = O.search_dense(oms, targets)
    assert es.max() > 200          # the embedded copies give real matches
    for kern in (BATCHED_KERNELS if n_masks >= 16 else ("auto",)):
        ctx.set_match_kernel(kern)
        scores, mirrored = ms.search_dense(lib)
        assert np.array_equal(scores, es), kern
        assert np.array_equal(mirrored, em), kern
        if xys <= 2:
            assert ctx.last_stats()["match_kernel"] == {"cand": 1, "band": 2, "auto": 3}[kern]
        elif xys == 4 and kern == "cand":
            assert ctx.last_stats()["match_kernel"] == 1
    ctx.set_match_kernel("auto")
    ms.close()


def test_wide_tolerance_uses_16_byte_records(ctx, synth):
    """pixColorFluctuation 90 makes intervals longer than the compact palette format holds: both batched kernels fall back to
    the 16-byte records."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    params = (20, 20, 0.9, 2, True)
    ms = _maskset(ctx, params, rects)
    ms.add_rgb(masks[:17])
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.9, 2, rects) for x in masks[:17]]
    es, em, _ = O.search_dense(oms, targets[:20])
    for kern in BATCHED_KERNELS:
        ctx.set_match_kernel(kern)
        scores, mirrored = ms.search_dense(lib)
        assert np.array_equal(scores[:, :20], es), kern
        assert np.array_equal(mirrored[:, :20], em), kern
    ctx.set_match_kernel("auto")
    ms.close()


def test_many_colour_classes_run_on_the_candidate_kernel(ctx, synth):
    """The reverse search: brightness-scaled LM images used as masks (any ImageArray is a legal query,
    ColorDepthSearchAlgorithmProvider.java:20-32).  A group of such masks has more colour classes than a shared-memory palette
    holds (2 047), so the word lists carry the rank intervals themselves (cds_cand.cuh: WIDE) and the candidate kernel still runs;
    the band kernel uses the 16-byte records for such a group."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    rng = np.random.default_rng(5)
    lm = (targets[:18].astype(np.float32) * rng.uniform(0.35, 1.0, size=(18, H, W, 1)).astype(np.float32)).astype(np.uint8)
    codes = _np_expected_codes(lm.reshape(-1, 3), 20)
    classes = np.unique(codes[(codes & 0x80000000) == 0] >> 8)
    assert len(classes) > 4096, len(classes)            # far more than one palette
    params = (20, 20, 0.01, 2, True)
    ms = _maskset(ctx, params, rects)
    sizes = ms.add_rgb(lm)
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in lm]
    assert sizes.tolist() == [m.size for m in oms]
    es, em, _ = O.search_dense(oms, targets[:24])
    assert es.max() > 1000          # a scaled image still matches its original within the tolerance in places
    for kern in BATCHED_KERNELS:
        ctx.set_match_kernel(kern)
        scores, mirrored = ms.search_dense(lib)
        assert ctx.last_stats()["match_kernel"] == {"cand": 1, "band": 2}[kern]
        assert np.array_equal(scores[:, :24], es), kern
        assert np.array_equal(mirrored[:, :24], em), kern
    ctx.set_match_kernel("auto")
    ms.close()
    # a set that mixes a many-class group with ordinary masks (one group here): still one consistent list format
    ms = _maskset(ctx, params, rects)
    mixed = np.concatenate([masks[:8], lm[:10]])
    ms.add_rgb(mixed)
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in mixed]
    es, em, _ = O.search_dense(oms, targets[:12])
    scores, mirrored = ms.search_dense(lib)
    assert ctx.last_stats()["match_kernel"] == 1
    assert np.array_equal(scores[:, :12], es) and np.array_equal(mirrored[:, :12], em)
    ms.close()
    # ordinary masks with the interval-carrying lists forced on ("wide_lists"), every supported xyShift: the palette lists' scores
    for xys in (0, 2, 4):
        got = []
        for wide in (0, 1):
            ctx.set_option("wide_lists", wide)
            ms = _maskset(ctx, (20, 20, 0.01, xys, True), rects)
            ms.add_rgb(masks)
            got.append(ms.search_dense(lib))
            assert ctx.last_stats()["match_kernel"] == 1
            ms.close()
        ctx.set_option("wide_lists", 0)
        assert np.array_equal(got[0][0], got[1][0]) and np.array_equal(got[0][1], got[1][1]), xys
        assert got[0][0].max() > 200


def test_topk_matches_sorted_dense(ctx, synth):
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    sizes = ms.add_rgb(masks)
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in masks]
    es, em, _ = O.search_dense(oms, targets)
    for k, pct in ((5, 0.0), (300, 0.0), (16, 1.0)):
        score, target, mirrored, count = ms.search_topk(lib, k, pct)
        for m in range(len(masks)):
            cand = [(-int(es[m, t]), t) for t in range(len(targets)) if O.is_match(es[m, t], es[m, t] / sizes[m], pct)]
            cand.sort()
            cand = cand[:k]
            assert count[m] == len(cand)
            assert score[m, :count[m]].tolist() == [-s for s, _ in cand]
            assert target[m, :count[m]].tolist() == [t for _, t in cand]
            assert mirrored[m, :count[m]].tolist() == [int(em[m, t]) for _, t in cand]
    ms.close()


@pytest.mark.parametrize("chunk", [256, 16, 7])
def test_stream_search_equals_library_search(ctx, synth, chunk):
    """cds_search_stream_rgb (host targets, chunked upload overlapping the search) == cds_library_add_rgb + cds_search_topk."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    ctx.set_option("stream_chunk", chunk)
    for n_masks, k, pct in ((24, 300, 0.0), (24, 5, 1.0), (3, 9, 0.0)):
        ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
        ms.add_rgb(masks[:n_masks])
        exp = ms.search_topk(lib, k, pct)
        got = ms.search_stream(targets, k, pct)
        assert np.array_equal(got[3], exp[3])
        for m in range(n_masks):
            c = exp[3][m]
            for a, b in zip(got[:3], exp[:3]):
                assert np.array_equal(a[m, :c], b[m, :c]), (chunk, n_masks, k, m)
        ms.close()
    # degenerate inputs
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    ms.add_rgb(masks[:2])
    s, t, mir, cnt = ms.search_stream(targets[:0], 4, 0.0)
    assert cnt.tolist() == [0, 0]
    ms.close()
    ctx.set_option("stream_chunk", 256)


@pytest.mark.parametrize("pct", [0.0, 1.0])
def test_all_matches_search(ctx, synth, pct):
    """cds_search_stream_matches_rgb returns EVERY pair passing ColorMIPSearch.isMatch (the reference keeps all of them,
    LocalColorMIPSearchProcessor.java:93-105), ordered by mask, descending score, ascending target."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    sizes = ms.add_rgb(masks)
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in masks]
    es, em, _ = O.search_dense(oms, targets)
    exp = []
    for m in range(len(masks)):
        row = [(m, t, int(es[m, t]), int(em[m, t])) for t in range(len(targets)) if O.is_match(es[m, t], es[m, t] / sizes[m], pct)]
        exp += sorted(row, key=lambda r: (-r[2], r[1]))
    ctx.set_option("stream_chunk", 16)
    got = ms.search_stream_matches(targets, pct)
    assert list(zip(got[0].tolist(), got[1].tolist(), got[2].tolist(), got[3].tolist())) == exp
    got = ms.search_matches(lib, pct)                       # the same over the resident library
    assert list(zip(got[0].tolist(), got[1].tolist(), got[2].tolist(), got[3].tolist())) == exp
    # too small a buffer: CDS_ERR_CAPACITY and the required size
    if len(exp) > 1:
        with pytest.raises(capi.CdsError) as e:
            ms.search_stream_matches(targets, pct, capacity=len(exp) - 1)
        assert e.value.status == capi.CDS_ERR_CAPACITY and str(len(exp)) in e.value.message
    ctx.set_option("stream_chunk", 256)
    ms.close()


def test_topk_with_occupancy_built_per_chunk(ctx, synth):
    """cds_search_topk when the occupancy bitmaps are not kept next to the planes (libraries near the memory limit)."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    ms.add_rgb(masks)
    exp = ms.search_topk(lib, 40, 1.0)
    ctx.set_option("resident_occupancy", 0)
    ctx.set_option("stream_chunk", 24)
    got = ms.search_topk(lib, 40, 1.0)
    assert ctx.last_stats()["match_kernel"] == 1 and ctx.last_stats()["chunked"] == 1
    ctx.set_option("resident_occupancy", 1)
    ctx.set_option("stream_chunk", 256)
    assert np.array_equal(got[3], exp[3])
    for m in range(len(masks)):
        c = exp[3][m]
        for a, b in zip(got[:3], exp[:3]):
            assert np.array_equal(a[m, :c], b[m, :c])
    ms.close()


def test_threshold_rebake_roundtrip(ctx, synth):
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    res = {}
    for dthr in (20, 120, 20):
        ms = _maskset(ctx, (20, dthr, 0.01, 2, True), rects)
        ms.add_rgb(masks[:3])
        res.setdefault(dthr, []).append(ms.search_dense(lib)[0])
        ms.close()
    assert np.array_equal(res[20][0], res[20][1])
    om = [O.PixelMatchMask(x, 20, True, 120, 0.01, 2, rects) for x in masks[:3]]
    assert np.array_equal(res[120][0], O.search_dense(om, targets)[0])


def test_device_generator_equals_host_generator(ctx):
    for kind in (0, 1):
        dev = ctx.synth_rgb(kind, 123, 3, 3, W, H, on_device=True)
        host = capi.synth_rgb_host(kind, 123, 3, 3, W, H)
        assert np.array_equal(dev, host)
    assert np.array_equal(ctx.synth_gradient(123, 3, 1, W, H, on_device=True), capi.synth_gradient_host(123, 3, 1, W, H))


def test_generated_library_equals_uploaded_library(ctx):
    rects = O.label_rects(W, H)
    masks = capi.synth_rgb_host(0, 9, 0, 17, W, H)
    lib_a = capi.Library(ctx, W, H, 70)
    lib_a.generate_synthetic(9, 100, 66)
    lib_b = capi.Library(ctx, W, H, 70)
    lib_b.add_rgb(capi.synth_rgb_host(1, 9, 100, 66, W, H))
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    ms.add_rgb(masks)
    a = ms.search_dense(lib_a)
    b = ms.search_dense(lib_b)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for x in (ms, lib_a, lib_b):
        x.close()


@pytest.mark.parametrize("shape", [(100, 50), (100, 300), (37, 23), (640, 480), (2048, 36), (257, 1024)])
def test_small_and_odd_image_sizes(ctx, shape):
    w, h = shape
    rng = np.random.default_rng(w * 1000 + h)
    lut = O.lut().astype(np.uint8)

    def rand_img(density):
        img = np.zeros((h, w, 3), np.uint8)
        sel = rng.random((h, w)) < density
        z = rng.integers(0, 256, (h, w))
        br = rng.integers(20, 256, (h, w, 1))
        col = (lut[z].astype(np.int64) * br // 255).astype(np.uint8)
        img[sel] = col[sel]
        return img

    masks = np.stack([rand_img(0.2) for _ in range(18)])
    targets = np.stack([rand_img(0.5) for _ in range(9)])
    rects = np.array([[0, 0, w // 3, h // 4]], np.int32)
    lib = capi.Library(ctx, w, h, 9)
    lib.add_rgb(targets)
    for params in ((20, 20, 0.02, 2, True), (20, 20, 0.02, 0, True), (20, 20, 0.02, 4, False)):
        mthr, dthr, ztol, xys, mirror = params
        for n, kern in ((18, "cand"), (18, "band"), (2, "auto")):
            ctx.set_match_kernel(kern)
            ms = capi.MaskSet(ctx, w, h, mthr, dthr, ztol, xys, mirror, rects)
            ms.add_rgb(masks[:n])
            oms = [O.PixelMatchMask(x, mthr, mirror, dthr, ztol, xys, rects) for x in masks[:n]]
            scores, mirrored = ms.search_dense(lib)
            es, em, _ = O.search_dense(oms, targets)
            assert np.array_equal(scores, es), (shape, params, n, kern)
            assert np.array_equal(mirrored, em), (shape, params, n, kern)
            ms.close()
    ctx.set_match_kernel("auto")
    lib.close()


def test_extreme_masks_in_a_batch(ctx, synth):
    """An empty mask, a mask that covers the whole image and a one-pixel mask next to ordinary ones, through every kernel."""
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    lut = O.lut().astype(np.uint8)
    full = np.zeros((H, W, 3), np.uint8)
    full[:] = lut[(np.arange(W)[None, :] // 5 + np.arange(H)[:, None] // 3) % 256]          # every pixel set, slowly varying slices
    one = np.zeros((H, W, 3), np.uint8)
    one[300, 700] = lut[100]
    batch = np.concatenate([np.zeros((1, H, W, 3), np.uint8), full[None], one[None], masks[:15]])
    tg = targets[:6]
    small = capi.Library(ctx, W, H, 6)
    small.add_rgb(tg)
    params = (20, 20, 0.01, 2, True)
    oms = [O.PixelMatchMask(x, 20, True, 20, 0.01, 2, rects) for x in batch]
    es, em, _ = O.search_dense(oms, tg)
    assert oms[0].size == 0 and oms[1].size > 600000 and oms[2].size == 1
    for kern in ("cand", "band", "gather"):
        ctx.set_match_kernel(kern)
        ms = _maskset(ctx, params, rects)
        sizes = ms.add_rgb(batch)
        assert sizes.tolist() == [m.size for m in oms]
        scores, mirrored = ms.search_dense(small)
        assert np.array_equal(scores, es), kern
        assert np.array_equal(mirrored, em), kern
        s, t, m, c = ms.search_topk(small, 4, 0.0)
        assert c[0] == 0                                   # an empty mask never matches
        ms.close()
    ctx.set_match_kernel("auto")
    small.close()


def test_concurrent_callers_on_one_context(ctx, synth):
    """The reference calls calculateMatchingScore on one shared algorithm instance from a pool of ~40 threads
    (LocalColorMIPSearchProcessor.java:93-105): calls on one cds_ctx from many host threads must be safe and correct."""
    from concurrent.futures import ThreadPoolExecutor
    masks, targets, lib = synth
    rects = O.label_rects(W, H)
    ms = _maskset(ctx, (20, 20, 0.01, 2, True), rects)
    sizes = ms.add_rgb(masks[:17])
    dense, dmir = ms.search_dense(lib)

    def pair(job):
        m, t = job
        return ms.score_pair(m, targets[t])

    jobs = [(m, t) for m in range(0, 17, 4) for t in range(0, 70, 7)]
    with ThreadPoolExecutor(8) as ex:
        futs = [ex.submit(pair, j) for j in jobs]
        tk = ex.submit(lambda: ms.search_topk(lib, 5, 0.0))
        res = [f.result() for f in futs]
        topk = tk.result()
    for (m, t), (score, ratio, mir) in zip(jobs, res):
        assert (score, mir) == (int(dense[m, t]), bool(dmir[m, t]))
        assert ratio == score / sizes[m]
    for m in range(17):
        order = sorted((j for j in range(70) if dense[m, j] > 0), key=lambda j: (-int(dense[m, j]), j))[:5]
        assert topk[1][m, :topk[3][m]].tolist() == order
    ms.close()


def test_two_contexts_on_one_device_run_concurrently(synth):
    """Each context owns its streams and kernel scratch: searches of two contexts on the same GPU may overlap."""
    from concurrent.futures import ThreadPoolExecutor
    masks, targets, _ = synth
    rects = O.label_rects(W, H)

    def run(seed_offset):
        c = capi.Context(n_dev=1)
        lib = capi.Library(c, W, H, 70)
        lib.add_rgb(targets[::-1] if seed_offset else targets)
        ms = capi.MaskSet(c, W, H, 20, 20, 0.01, 2, True, rects)
        ms.add_rgb(masks)
        outs = [ms.search_dense(lib)[0] for _ in range(6)]
        ms.close(); lib.close(); c.close()
        return outs

    with ThreadPoolExecutor(2) as ex:
        a, b = ex.map(run, (0, 1))
    for x in a[1:]:
        assert np.array_equal(x, a[0])
    for x in b:
        assert np.array_equal(x[:, ::-1], a[0])


def test_edge_cases_and_errors(ctx, fixtures):
    rects = O.label_rects(W, H)
    # odd xyShift -> IllegalArgumentException (ColorDepthSearchAlgorithmProviderFactory.java:57-60)
    with pytest.raises(capi.CdsIllegalArgument) as e:
        capi.MaskSet(ctx, W, H, 20, 20, 0.01, 3, True, rects)
    assert "even number" in e.value.message
    # empty mask -> score 0, ratio 0, not mirrored; checked BEFORE the size test (PixelMatch...:169-175)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    sizes = ms.add_rgb(np.stack([np.zeros((H, W, 3), np.uint8), fixtures["em_LPLC2"]]))
    assert sizes.tolist() == [0, 1897]
    assert ms.score_pair(0, fixtures["lm_GMR"]) == (0, 0.0, False)
    assert ms.score_pair(0, np.zeros((10, 10, 3), np.uint8)) == (0, 0.0, False)
    with pytest.raises(capi.CdsIllegalArgument) as e:
        ms.score_pair(1, np.zeros((10, 10, 3), np.uint8))
    assert e.value.status == capi.CDS_ERR_SIZE_MISMATCH and "Invalid image size" in e.value.message
    # size mismatch between mask set and library
    lib = capi.Library(ctx, 100, 50, 4)
    lib.add_rgb(np.zeros((1, 50, 100, 3), np.uint8))
    with pytest.raises(capi.CdsIllegalArgument):
        ms.search_dense(lib)
    # empty library / capacity
    lib2 = capi.Library(ctx, W, H, 2)
    s, m = ms.search_dense(lib2)
    assert s.shape == (2, 0)
    lib2.add_rgb(np.stack([fixtures["lm_GMR"], fixtures["lm_BJD"]]))
    with pytest.raises(capi.CdsError) as e:
        lib2.add_rgb(fixtures["lm_GMR"])
    assert e.value.status == capi.CDS_ERR_CAPACITY
    s, m = ms.search_dense(lib2)
    assert s[0].tolist() == [0, 0] and s[1, 0] > 0
    score, target, mirrored, count = ms.search_topk(lib2, 4, 0.0)
    assert count.tolist()[0] == 0 and count[1] >= 1
    for x in (ms, lib, lib2):
        x.close()
