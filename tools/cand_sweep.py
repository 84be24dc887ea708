"""Sweeps the candidate kernel's tuning knobs (cds_ctx_set_option cand_*) on BASELINE configs[1]'s parameters in one process:
one resident synthetic library, one prepared mask set, a few timed searches per setting.  Prints one JSON line per setting.

    python tools/cand_sweep.py [--masks 1000] [--targets 4096] [--reps 3]
"""
import argparse, json, sys
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O

SEED = 0xC0FFEE
ap = argparse.ArgumentParser()
ap.add_argument("--masks", type=int, default=1000)
ap.add_argument("--targets", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--width", type=int, default=1210)
ap.add_argument("--height", type=int, default=566)
ap.add_argument("--settings", default="wait=0;wait=1;wait=32;wait=64;wait=100;wait=200;wait=400;wait=0,hint=1")
ap.add_argument("--wide", type=int, default=0, help="1: word lists that carry the rank intervals (cds_ctx_set_option wide_lists)")
a = ap.parse_args()
W, H = a.width, a.height
rects = O.label_rects(W, H)
ctx = capi.Context(device_ids=[0])
lib = capi.Library(ctx, W, H, a.targets)
lib.generate_synthetic(SEED, 0, a.targets)
masks = np.concatenate([ctx.synth_rgb(0, SEED, i, min(64, a.masks - i), W, H, on_device=True) for i in range(0, a.masks, 64)])
ctx.set_option("wide_lists", a.wide)
ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
ms.add_rgb(masks)
base = None
for setting in a.settings.split(";"):
    kv = dict(x.split("=") for x in setting.split(","))
    ctx.set_option("cand_wait_mode", int(kv.get("wait", 0)))
    ctx.set_option("cand_l2_hint", int(kv.get("hint", 0)))
    ctx.set_option("cand_warps", int(kv.get("warps", 31)))
    ctx.set_option("cand_stages", int(kv.get("stages", 2)))
    ctx.set_option("cand_max_rows", int(kv.get("rows", 0)))
    res = ms.search_topk(lib, 300, 1.0)
    ms_total = 0.0
    for _ in range(a.reps):
        res = ms.search_topk(lib, 300, 1.0)
        ms_total += ctx.last_stats()["match_kernel_ms"]
    same = True
    if base is None:
        base = res
    else:
        same = bool(np.array_equal(res[3], base[3]) and np.array_equal(res[0], base[0]) and np.array_equal(res[1], base[1]))
    print(json.dumps({"setting": setting, "wide_lists": a.wide, "comparisons_per_s": a.masks * a.targets * a.reps / (ms_total * 1e-3), "ms_per_search": ms_total / a.reps,
                      "same_result_as_first": same}), flush=True)
