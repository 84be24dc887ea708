"""The N > 1 host path (target shards, masks replicated, host merge of per-shard top-K) on CPU: two gloo ranks, scores from
the oracle, merged result == single-process top-K."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from colormipsearch_b200 import sharding

W, H = 320, 200
K = 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_data():
    from oracle import oracle as O
    rng = np.random.default_rng(42)
    lut = O.lut().astype(np.uint8)

    def img(density):
        a = np.zeros((H, W, 3), np.uint8)
        sel = rng.random((H, W)) < density
        a[sel] = lut[rng.integers(0, 256, (H, W))][sel]
        return a

    masks = [img(0.05) for _ in range(5)]
    targets = np.stack([img(0.3) for _ in range(23)])      # 23: shards of unequal size
    return masks, targets


def _topk_from_dense(scores, mirrored, sizes, k, pct):
    from oracle import oracle as O
    M, T = scores.shape
    out = (np.zeros((M, k), np.int32), np.full((M, k), -1, np.int64), np.zeros((M, k), np.uint8), np.zeros(M, np.int32))
    for m in range(M):
        cand = sorted((-int(scores[m, t]), t) for t in range(T) if O.is_match(scores[m, t], scores[m, t] / max(sizes[m], 1), pct))[:k]
        out[3][m] = len(cand)
        for i, (s, t) in enumerate(cand):
            out[0][m, i], out[1][m, i], out[2][m, i] = -s, t, mirrored[m, t]
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    masks, targets = _make_data()
    rects = np.zeros((0, 4), np.int32)
    oms = [O.PixelMatchMask(m, 20, True, 20, 0.02, 2, rects) for m in masks]
    lo, hi = sharding.shard_range(rank, world, len(targets))
    s, mir, _ = O.search_dense(oms, targets[lo:hi], 1)
    local = _topk_from_dense(s, mir, [m.size for m in oms], K, 0.5)
    merged = sharding.gather_and_merge_topk(local, K, lo)
    if rank == 0:
        q.put(merged)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_merge_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from oracle import oracle as O
    masks, targets = _make_data()
    rects = np.zeros((0, 4), np.int32)
    oms = [O.PixelMatchMask(m, 20, True, 20, 0.02, 2, rects) for m in masks]
    s, mir, _ = O.search_dense(oms, targets, 1)
    exp = _topk_from_dense(s, mir, [m.size for m in oms], K, 0.5)
    for a, b in zip(merged, exp):
        assert np.array_equal(a, b)
    assert exp[3].max() > 0


def test_shard_ranges_cover_everything():
    for T in (0, 1, 7, 100000):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(i, world, T) for i in range(world)]
            assert r[0][0] == 0 and r[-1][1] == T
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))


def test_merge_ties_by_ascending_target():
    a = (np.array([[9, 5, 5]], np.int32), np.array([[4, 1, 8]], np.int64), np.zeros((1, 3), np.uint8), np.array([3], np.int32))
    b = (np.array([[9, 5, 0]], np.int32), np.array([[2, 3, -1]], np.int64), np.ones((1, 3), np.uint8), np.array([2], np.int32))
    s, t, r, c = sharding.merge_topk([a, b], 4)
    assert s[0].tolist() == [9, 9, 5, 5] and t[0].tolist() == [2, 4, 1, 3] and c[0] == 4
    assert r[0].tolist() == [1, 0, 0, 1]
