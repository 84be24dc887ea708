"""BASELINE configs[3]: "xyShift=4 with mirrored masks and tight pixColorFluctuation" (17 distinct offsets x 2 orientations = 34
variants, zTolerance 0.005; the Java reference throws for xyShift >= 4, so the oracle is the specification).  Resident search of
`--masks` x `--targets`, top-300; a sample of cells is checked against the oracle.  Prints one JSON line."""
import argparse, json, sys, time
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O

W, H, SEED = 1210, 566, 0xC0FFEE
ap = argparse.ArgumentParser()
ap.add_argument("--masks", type=int, default=1000)
ap.add_argument("--targets", type=int, default=10000)
a = ap.parse_args()
M, T = a.masks, a.targets
rects = O.label_rects(W, H)
ctx = capi.Context(device_ids=[0])
lib = capi.Library(ctx, W, H, T)
lib.generate_synthetic(SEED, 0, T)
masks = np.concatenate([ctx.synth_rgb(0, SEED, i, min(64, M - i), W, H, on_device=True) for i in range(0, M, 64)])
ms = capi.MaskSet(ctx, W, H, 20, 20, 0.005, 4, True, rects)
sizes = ms.add_rgb(masks)
for _ in range(3):
    res = ms.search_topk(lib, 300, 1.0)
dev_ms = 0.0
steps = 5
for _ in range(steps):
    res = ms.search_topk(lib, 300, 1.0)
    dev_ms += ctx.last_stats()["total_device_ms"]
st = ctx.last_stats()
rng = np.random.default_rng(4)
pick_m = np.sort(rng.choice(M, 12, replace=False)); pick_t = np.sort(rng.choice(T, 24, replace=False))
sub = capi.MaskSet(ctx, W, H, 20, 20, 0.005, 4, True, rects)
sub.add_rgb(masks[pick_m])
small = capi.Library(ctx, W, H, len(pick_t))
tg = np.stack([ctx.synth_rgb(1, SEED, int(j), 1, W, H, on_device=True)[0] for j in pick_t])
small.add_rgb(tg)
ctx.set_match_kernel("cand")
sub_ms = capi.MaskSet(ctx, W, H, 20, 20, 0.005, 4, True, rects)
sub_ms.add_rgb(np.concatenate([masks[pick_m], masks[:8]]))          # >= 16 masks: the batched kernel
dense, dmir = sub_ms.search_dense(small)
oms = [O.PixelMatchMask(masks[i], 20, True, 20, 0.005, 4, rects) for i in pick_m]
es, em, _ = O.search_dense(oms, tg)
ok = bool(np.array_equal(dense[:12], es) and np.array_equal(dmir[:12], em))
print(json.dumps({"workload": "BASELINE configs[3]: %d masks x %d targets, xyShift 4 (34 variants), zTol 0.005, mirror, top-300" % (M, T),
                  "comparisons_per_s": M * T * steps / (dev_ms * 1e-3), "ms_per_search": dev_ms / steps, "match_kernel": st["match_kernel"],
                  "parity_sampled_cells": ok, "matches_returned": int(res[3].sum())}))
