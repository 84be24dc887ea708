"""The single-pair call behind the micro-batching queue (cds_pairq_*): what ~40 pool threads of the reference do with
ColorDepthSearchAlgorithm.calculateMatchingScore (LocalColorMIPSearchProcessor.java:93-105).  Scores must be exactly the batched
search's (which the other tests pin on the oracle), whatever the interleaving, the batch size and the cache size."""
import threading

import numpy as np
import pytest

from colormipsearch_b200 import capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

W, H = 1210, 566


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(n_dev=1)
    yield c
    c.close()


@pytest.fixture(scope="module")
def data():
    masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 6, W, H)
    targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 12, W, H)
    return masks, targets


@pytest.mark.parametrize("params", [(20, 20, 0.01, 2, True), (100, 100, 0.02, 0, True), (20, 20, 0.005, 4, False)], ids=["production", "defaults", "xy4_nomirror"])
def test_queue_scores_equal_dense_search_and_oracle(ctx, data, params):
    masks, targets = data
    mthr, dthr, ztol, xys, mirror = params
    rects = O.label_rects(W, H)
    ms = capi.MaskSet(ctx, W, H, mthr, dthr, ztol, xys, mirror, rects)
    sizes = ms.add_rgb(masks)
    lib = capi.Library(ctx, W, H, 16)
    lib.add_rgb(targets)
    dense, dmir = ms.search_dense(lib)
    lib.close()
    rng = np.random.default_rng(4)
    pm = np.repeat(np.arange(6), 12).astype(np.int32)
    pt = np.tile(np.arange(12), 6).astype(np.int64)
    perm = rng.permutation(len(pm))
    pm, pt = pm[perm], pt[perm]
    keys = (np.arange(12) + 1000).astype(np.uint64)
    for max_batch, cache, use_keys, threads in ((64, 256, True, 8), (4, 8, True, 16), (16, 64, False, 5), (1, 2, True, 3)):
        q = capi.PairQueue(ctx, ms, max_batch=max_batch, max_wait_us=100, cache_targets=cache)
        sc, mir, secs = q.drive(targets, keys if use_keys else None, pm, pt, threads)
        st = q.stats()
        assert np.array_equal(sc, dense[pm, pt]), (max_batch, cache, use_keys)
        assert np.array_equal(mir, dmir[pm, pt].astype(bool))
        assert st["requests"] == len(pm) and 1 <= st["batches"] <= len(pm)
        if use_keys and cache >= 12:
            assert st["uploads"] == 12                       # every target crossed PCIe once, whatever the interleaving
        if not use_keys:
            assert st["uploads"] == len(pm)
        q.close()
    # the oracle on a few pairs (the dense search itself is pinned on it elsewhere)
    for i in range(0, len(pm), 17):
        om = O.PixelMatchMask(masks[pm[i]], mthr, mirror, dthr, ztol, xys, rects)
        s, ratio, m = om.score(targets[pt[i]])
        assert (s, m) == (int(dense[pm[i], pt[i]]), bool(dmir[pm[i], pt[i]]))
    ms.close()


def test_queue_from_python_threads_and_errors(ctx, data):
    masks, targets = data
    rects = O.label_rects(W, H)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    empty = np.zeros((1, H, W, 3), np.uint8)
    sizes = ms.add_rgb(np.concatenate([masks[:3], empty]))
    lib = capi.Library(ctx, W, H, 16)
    lib.add_rgb(targets)
    dense, dmir = ms.search_dense(lib)
    lib.close()
    q = capi.PairQueue(ctx, ms, max_batch=8, max_wait_us=200, cache_targets=32)
    out = {}

    def work(tid):
        for j in range(12):
            m = (tid + j) % 3
            out[(tid, j)] = (m, j, q.score(m, targets[j], key=j + 1))

    th = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for (tid, j), (m, t, (s, ratio, mir)) in out.items():
        assert (s, mir) == (int(dense[m, t]), bool(dmir[m, t]))
        assert ratio == s / sizes[m]
    assert q.score(3, targets[0], key=1) == (0, 0.0, False)          # empty mask: (0, 0, false) before anything else (:169-170)
    with pytest.raises(capi.CdsIllegalArgument) as e:
        q.score(0, np.zeros((10, 10, 3), np.uint8))
    assert e.value.status == capi.CDS_ERR_SIZE_MISMATCH and "Invalid image size" in str(e.value)
    with pytest.raises(capi.CdsIllegalArgument):
        q.score(99, targets[0])
    # the same key must mean the same pixels: a DIFFERENT key for other pixels gives the other score
    a = q.score(0, targets[5], key=777)
    b = q.score(0, targets[6], key=778)
    assert (a[0], b[0]) == (int(dense[0, 5]), int(dense[0, 6]))
    q.close()
    ms.close()


def test_masks_added_after_the_queue_was_created(ctx, data):
    """One provider = one mask set + one queue, an algorithm per mask: masks keep arriving while pairs are being scored."""
    masks, targets = data
    rects = O.label_rects(W, H)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb(masks[:2])
    q = capi.PairQueue(ctx, ms, max_batch=8, max_wait_us=100, cache_targets=32)
    first = [q.score(m, targets[t], key=t + 1) for m in range(2) for t in range(4)]
    with pytest.raises(capi.CdsIllegalArgument):
        q.score(2, targets[0], key=1)                      # not there yet
    ms.add_rgb(masks[2:6])
    keys = (np.arange(12) + 1).astype(np.uint64)
    pm = np.repeat(np.arange(6), 12).astype(np.int32)
    pt = np.tile(np.arange(12), 6).astype(np.int64)
    sc, mir, _ = q.drive(targets, keys, pm, pt, 12)
    lib = capi.Library(ctx, W, H, 16)
    lib.add_rgb(targets)
    dense, dmir = ms.search_dense(lib)
    lib.close()
    assert np.array_equal(sc, dense[pm, pt]) and np.array_equal(mir, dmir[pm, pt].astype(bool))
    assert [f[0] for f in first] == [int(dense[m, t]) for m in range(2) for t in range(4)]
    q.close()
    ms.close()
