"""BASELINE.json configurations at (or near) their real sizes, checked through size-independent properties and oracle samples.

configs[0]  1 mask x 1,000 targets, CLI defaults (thr 100/100, pixColorFluctuation 2, xyShift 0, mirror): every cell vs the oracle.
configs[1]  1,000 masks x N targets, production parameters, top-300: ordering / floor / count properties on all masks, every
            cell of 24 sampled masks against the oracle on sampled targets, top-K of those masks against their sorted dense rows,
            streaming search == resident search.
configs[2]  gradientScores: the top-300 pixel matches of 1,000 masks feed the shape score; >= 256 pairs (gap, high expression
            area, mirrored flag) and the normalised scores of their masks against the oracle.
configs[3]  xyShift 4 + pixColorFluctuation 0.5 (the Java reference throws here; the oracle is the specification).
configs[4]  the full-library sweep in miniature: several mask groups (3,000 masks) x 4,096 targets with the occupancy bitmaps built
            per target chunk (the mode a 50,000-target shard runs in), FULL top-300 lists of 16 masks against the oracle.
"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

W, H = 1210, 566
SEED = 0xC0FFEE


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


def _synth(ctx, kind, first, n):
    return np.concatenate([ctx.synth_rgb(kind, SEED, first + i, min(64, n - i), W, H, on_device=True) for i in range(0, n, 64)])


def test_config0_one_mask_thousand_targets(ctx):
    rects = O.label_rects(W, H)
    mask = _synth(ctx, 0, 7, 1)
    lib = capi.Library(ctx, W, H, 1000)
    lib.generate_synthetic(SEED, 0, 1000)
    targets = _synth(ctx, 1, 0, 1000)
    ms = capi.MaskSet(ctx, W, H, 100, 100, 0.02, 0, True, rects)
    ms.add_rgb(mask)
    scores, mirrored = ms.search_dense(lib)
    om = [O.PixelMatchMask(mask[0], 100, True, 100, 0.02, 0, rects)]
    es, em, _ = O.search_dense(om, targets)
    assert np.array_equal(scores, es) and np.array_equal(mirrored, em)
    s, t, m, c = ms.search_topk(lib, 300, 1.0)
    order = sorted((j for j in range(1000) if O.is_match(es[0, j], es[0, j] / om[0].size, 1.0)), key=lambda j: (-int(es[0, j]), j))[:300]
    assert c[0] == len(order) and t[0, :c[0]].tolist() == order
    ms.close()
    lib.close()


def _check_topk_properties(score, target, mirrored, count, sizes, k, pct, n_targets):
    for m in range(len(sizes)):
        c = int(count[m])
        assert 0 <= c <= k
        s, t = score[m, :c].astype(np.int64), target[m, :c]
        assert np.all((t >= 0) & (t < n_targets)) and len(set(t.tolist())) == c
        # descending score, ties by ascending target index
        assert np.all((s[:-1] > s[1:]) | ((s[:-1] == s[1:]) & (t[:-1] < t[1:])))
        # ColorMIPSearch.isMatch floor
        assert all(O.is_match(int(x), int(x) / sizes[m], pct) for x in (s[-1:] if c else []))


@pytest.mark.parametrize("params,n_masks,n_targets", [
    ((20, 20, 0.01, 2, True), 1000, 2048),      # configs[1] parameters, one GPU's masks, a slice of its targets
    ((20, 20, 0.005, 4, True), 256, 1024),      # configs[3]
], ids=["config1", "config3_xy4"])
def test_batched_search_properties_and_samples(ctx, params, n_masks, n_targets):
    mthr, dthr, ztol, xys, mirror = params
    rects = O.label_rects(W, H)
    K, pct = 300, 1.0
    masks = _synth(ctx, 0, 0, n_masks)
    lib = capi.Library(ctx, W, H, n_targets)
    lib.generate_synthetic(SEED, 0, n_targets)
    ms = capi.MaskSet(ctx, W, H, mthr, dthr, ztol, xys, mirror, rects)
    sizes = ms.add_rgb(masks)
    res = ms.search_topk(lib, K, pct)
    assert ctx.last_stats()["match_kernel"] == 1                      # the candidate kernel
    _check_topk_properties(*res, sizes, K, pct, n_targets)
    assert int(res[3].sum()) > 0

    # sampled masks: dense rows vs oracle on sampled targets, and top-K vs the sorted dense row
    rng = np.random.default_rng(5)
    pick_m = np.sort(rng.choice(n_masks, 24, replace=False))
    pick_t = np.sort(rng.choice(n_targets, 40, replace=False))
    sub = capi.MaskSet(ctx, W, H, mthr, dthr, ztol, xys, mirror, rects)
    sub.add_rgb(masks[pick_m])
    dense, dmir = sub.search_dense(lib)
    sub.close()
    tg = np.stack([ctx.synth_rgb(1, SEED, int(j), 1, W, H, on_device=True)[0] for j in pick_t])
    oms = [O.PixelMatchMask(masks[i], mthr, mirror, dthr, ztol, xys, rects) for i in pick_m]
    es, em, _ = O.search_dense(oms, tg)
    assert np.array_equal(dense[:, pick_t], es)
    assert np.array_equal(dmir[:, pick_t], em)
    for a, i in enumerate(pick_m):
        row = dense[a].astype(np.int64)
        cand = [j for j in np.argsort(-row, kind="stable") if O.is_match(int(row[j]), row[j] / sizes[i], pct)][:K]
        c = int(res[3][i])
        assert c == len(cand)
        assert res[1][i, :c].tolist() == [int(j) for j in cand]
        assert res[0][i, :c].tolist() == [int(row[j]) for j in cand]
        assert res[2][i, :c].tolist() == [int(dmir[a, j]) for j in cand]

    # the streaming search over the same targets held on the host
    host_targets = _synth(ctx, 1, 0, n_targets)
    ctx.set_option("stream_chunk", 192)
    st = ms.search_stream(host_targets, K, pct)
    ctx.set_option("stream_chunk", 256)
    assert np.array_equal(st[3], res[3])
    for m in range(n_masks):
        c = int(res[3][m])
        assert np.array_equal(st[0][m, :c], res[0][m, :c]) and np.array_equal(st[1][m, :c], res[1][m, :c]) and np.array_equal(st[2][m, :c], res[2][m, :c])
    ms.close()
    lib.close()


def _oracle_topk(es_row, size, k, pct):
    """The reference's list for one mask: every isMatch target, descending matchingPixels, ties by ascending index."""
    cand = [j for j in np.argsort(-es_row.astype(np.int64), kind="stable") if O.is_match(int(es_row[j]), es_row[j] / size, pct)]
    return cand[:k]


def test_config1_full_topk_lists_vs_oracle(ctx):
    """configs[1] parameters: the complete top-300 lists of 16 masks (scores, order, mirrored flags) against oracle rows over ALL targets."""
    rects = O.label_rects(W, H)
    n_masks, n_targets, K, pct = 1000, 2048, 300, 1.0
    masks = _synth(ctx, 0, 0, n_masks)
    lib = capi.Library(ctx, W, H, n_targets)
    lib.generate_synthetic(SEED, 0, n_targets)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    sizes = ms.add_rgb(masks)
    res = ms.search_topk(lib, K, pct)
    assert ctx.last_stats()["match_kernel"] == 1
    pick = np.sort(np.random.default_rng(17).choice(n_masks, 16, replace=False))
    targets = _synth(ctx, 1, 0, n_targets)
    oms = [O.PixelMatchMask(masks[i], 20, True, 20, 0.01, 2, rects) for i in pick]
    es, em, _ = O.search_dense(oms, targets)
    for a, i in enumerate(pick):
        exp = _oracle_topk(es[a], sizes[i], K, pct)
        c = int(res[3][i])
        assert c == len(exp)
        assert res[1][i, :c].tolist() == [int(j) for j in exp]
        assert res[0][i, :c].tolist() == [int(es[a, j]) for j in exp]
        assert res[2][i, :c].tolist() == [int(em[a, j]) for j in exp]
    ms.close()
    lib.close()


def test_config4_sweep_many_groups_chunked_occupancy(ctx):
    rects = O.label_rects(W, H)
    n_masks, n_targets, K, pct = 3000, 4096, 300, 1.0
    lib = capi.Library(ctx, W, H, n_targets)
    lib.generate_synthetic(SEED, 0, n_targets)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    pick = np.sort(np.random.default_rng(23).choice(n_masks, 16, replace=False))
    keep = {}
    for i in range(0, n_masks, 64):
        n = min(64, n_masks - i)
        block = ctx.synth_rgb(0, SEED, i, n, W, H, on_device=True)
        ms.add_rgb(block)
        for m in pick[(pick >= i) & (pick < i + n)]:
            keep[int(m)] = block[m - i].copy()
    sizes = ms.sizes()
    ctx.set_option("resident_occupancy", 0)            # what a 50,000-target shard of configs[4] does: bitmaps per target chunk
    try:
        res = ms.search_topk(lib, K, pct)
        st = ctx.last_stats()
    finally:
        ctx.set_option("resident_occupancy", 1)
    assert st["match_kernel"] == 1 and st["chunked"] == 1
    _check_topk_properties(*res, sizes, K, pct, n_targets)
    res2 = ms.search_topk(lib, K, pct)                  # resident bitmaps: the same lists
    assert np.array_equal(res[3], res2[3])
    for m in range(0, n_masks, 37):
        c = int(res[3][m])
        assert all(np.array_equal(res[i][m, :c], res2[i][m, :c]) for i in range(3))
    targets = _synth(ctx, 1, 0, n_targets)
    oms = [O.PixelMatchMask(keep[int(i)], 20, True, 20, 0.01, 2, rects) for i in pick]
    es, em, _ = O.search_dense(oms, targets)
    for a, i in enumerate(pick):
        exp = _oracle_topk(es[a], sizes[i], K, pct)
        c = int(res[3][i])
        assert c == len(exp)
        assert res[1][i, :c].tolist() == [int(j) for j in exp]
        assert res[0][i, :c].tolist() == [int(es[a, j]) for j in exp]
        assert res[2][i, :c].tolist() == [int(em[a, j]) for j in exp]
    ms.close()
    lib.close()


def test_config2_gradient_scores(ctx):
    """colorDepthSearch -> gradientScores: top-300 pixel matches of 1,000 masks, shape score of every pair, normalised scores."""
    rects = O.label_rects(W, H)
    n_masks, n_targets, K, pct = 1000, 2048, 300, 1.0
    masks = _synth(ctx, 0, 0, n_masks)
    lib = capi.Library(ctx, W, H, n_targets)
    lib.generate_synthetic(SEED, 0, n_targets)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb(masks)
    score, target, mirrored, count = ms.search_topk(lib, K, pct)
    ms.close()
    lib.close()
    pair_mask = np.repeat(np.arange(n_masks, dtype=np.int32), count)
    pair_target = np.concatenate([target[m, :count[m]] for m in range(n_masks)]).astype(np.int64)
    pix = np.concatenate([score[m, :count[m]] for m in range(n_masks)]).astype(np.int32)
    assert len(pair_mask) >= 2000
    targets = _synth(ctx, 1, 0, n_targets)
    grads = np.concatenate([ctx.synth_gradient(SEED, i, min(64, n_targets - i), W, H, on_device=True) for i in range(0, n_targets, 64)])
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    qm, he_sizes = sms.add_rgb(masks)
    gap, he, mir = sms.score_pairs(targets, grads, None, pair_mask, pair_target)
    assert np.all(gap >= 0) and np.all(he >= 0)
    # the same pairs with the targets as TIFF files, a different call shape (pairs shuffled): the same numbers
    perm = np.random.default_rng(2).permutation(len(pair_mask))
    files = [capi.tiff_encode_rgb(t, 8, 32773) for t in targets]
    gap2, he2, mir2 = sms.score_pairs_tiff(files, grads, None, pair_mask[perm], pair_target[perm])
    assert np.array_equal(gap2, gap[perm]) and np.array_equal(he2, he[perm]) and np.array_equal(mir2, mir[perm])
    # oracle: every pair of 16 masks that have matches (>= 256 pairs in all)
    rng = np.random.default_rng(9)
    cand = [m for m in rng.permutation(n_masks) if count[m] >= 8]
    chosen, total = [], 0
    for m in cand:
        chosen.append(int(m)); total += min(int(count[m]), 24)
        if len(chosen) >= 16 and total >= 256:
            break
    assert total >= 256
    jobs = []
    for m in chosen:
        idx = np.nonzero(pair_mask == m)[0][:24]
        jobs += [(m, int(i)) for i in idx]
    with ThreadPoolExecutor(O.num_threads()) as ex:
        oms = dict(zip(chosen, ex.map(lambda m: O.ShapeMask(masks[m], 20, True, rects), chosen)))
        need_t = sorted({int(pair_target[i]) for _, i in jobs})
        zg = dict(zip(need_t, ex.map(lambda t: O.make_zgap(targets[t], 20, rects), need_t)))
        exp = list(ex.map(lambda j: oms[j[0]].score(targets[int(pair_target[j[1]])], grads[int(pair_target[j[1]])], zg[int(pair_target[j[1]])]), jobs))
    for (m, i), e in zip(jobs, exp):
        assert (int(gap[i]), int(he[i]), bool(mir[i])) == e, (m, i)
    for m in chosen:
        assert (int(qm[m]), int(he_sizes[m])) == (int(oms[m].qm.sum()), int(oms[m].he.sum()))
    # normalised scores per mask (CalculateGradientScoresCmd.normalizeScores), 1e-6 relative
    for m in chosen[:8]:
        sel = pair_mask == m
        got = capi.normalize_scores(pix[sel], gap[sel], he[sel])
        shape = [O.shape_score_2d(g, h) for g, h in zip(gap[sel], he[sel])]
        mx_p, mx_s = int(pix[sel].max()), max(shape)
        want = np.array([np.float32(O.normalized_score(int(p), s, mx_p, mx_s)) for p, s in zip(pix[sel], shape)])
        assert np.allclose(got, want, rtol=1e-6, atol=0)
    sms.close()
