"""BASELINE.json configurations at (or near) their real sizes, checked through size-independent properties and oracle samples.

configs[0]  1 mask x 1,000 targets, CLI defaults (thr 100/100, pixColorFluctuation 2, xyShift 0, mirror): every cell vs the oracle.
configs[1]  1,000 masks x N targets, production parameters, top-300: ordering / floor / count properties on all masks, every
            cell of 24 sampled masks against the oracle on sampled targets, top-K of those masks against their sorted dense rows,
            streaming search == resident search.
configs[3]  xyShift 4 + pixColorFluctuation 0.5 (the Java reference throws here; the oracle is the specification).
"""
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
