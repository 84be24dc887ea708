"""Host-side logic of the multi-GPU / multi-process path: contiguous target shards, masks replicated, per-shard top-K lists
merged on the host.  There is no data-path collective -- every (mask, target) comparison is independent -- so
the only exchange is a gather of M x K (score, target, mirrored) tuples, done with plain objects over the process group.

Ordering is the reference's: descending matchingPixels (colormipsearch-tools/.../cmd/ColorDepthSearchCmd.java:403-409), ties by
ascending global target index (what a single-threaded LocalColorMIPSearchProcessor run followed by the stable sort produces).
"""
import numpy as np


def shard_range(rank, world, n_targets):
    """Contiguous shard [lo, hi) of rank `rank` (shard g = [g*T/G, (g+1)*T/G))."""
    lo = rank * n_targets // world
    hi = (rank + 1) * n_targets // world
    return lo, hi


def merge_topk(parts, k):
    """parts: list (one per shard) of (score [M,K], target [M,K] GLOBAL indices, mirrored [M,K], count [M]).
    Returns the same 4-tuple for the union, k best per mask."""
    M = parts[0][0].shape[0]
    score = np.zeros((M, k), np.int32)
    target = np.full((M, k), -1, np.int64)
    mirrored = np.zeros((M, k), np.uint8)
    count = np.zeros(M, np.int32)
    for m in range(M):
        s = np.concatenate([p[0][m, :p[3][m]] for p in parts])
        t = np.concatenate([p[1][m, :p[3][m]] for p in parts])
        r = np.concatenate([p[2][m, :p[3][m]] for p in parts])
        order = np.lexsort((t, -s.astype(np.int64)))[:k]
        n = len(order)
        count[m] = n
        score[m, :n] = s[order]
        target[m, :n] = t[order]
        mirrored[m, :n] = r[order]
    return score, target, mirrored, count


def gather_and_merge_topk(local, k, first_target, group=None):
    """Every rank passes its local top-K (target indices LOCAL to its shard) and the global index of its first target;
    rank 0 receives the merged global top-K, other ranks None.  Works on any backend (gloo on CPU, nccl)."""
    import torch.distributed as dist
    score, target, mirrored, count = local
    glob = (score, np.where(target >= 0, target + first_target, -1), mirrored, count)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(glob, gathered, dst=0, group=group)
    if rank != 0:
        return None
    return merge_topk(gathered, k)
