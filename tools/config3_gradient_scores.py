"""BASELINE configs[2] at full size: "gradientScores: shape/area-gap scoring of the top-300 matches per mask with gradient + zgap
images for 1,000 masks" = 300,000 pairs.  Pixel-match top-300 per mask over 12,500 resident targets picks the
pairs; targets and gradient images then come from pinned host memory (zgap images derived on the device, as the reference's tests
do); a sample of pairs is checked against the oracle.  Prints one JSON line.

    python tools/config3_gradient_scores.py [--masks 1000] [--targets 12500]
"""
import argparse, json, sys, time
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O

W, H, SEED = 1210, 566, 0xC0FFEE
ap = argparse.ArgumentParser()
ap.add_argument("--masks", type=int, default=1000)
ap.add_argument("--targets", type=int, default=12500)
ap.add_argument("--check-pairs", type=int, default=32)
a = ap.parse_args()
M, T, K = a.masks, a.targets, 300
rects = O.label_rects(W, H)
ctx = capi.Context(device_ids=[0])
img = 3 * W * H

# pixel-match search picks the pairs (colorDepthSearch output feeding gradientScores)
lib = capi.Library(ctx, W, H, T)
lib.generate_synthetic(SEED, 0, T)
masks = np.concatenate([ctx.synth_rgb(0, SEED, i, min(64, M - i), W, H, on_device=True) for i in range(0, M, 64)])
ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
ms.add_rgb(masks)
score, target, mirrored, count = ms.search_topk(lib, K, 1.0)
ms.close(); lib.close()
pair_mask = np.repeat(np.arange(M, dtype=np.int32), count)
pair_target = np.concatenate([target[m, :count[m]] for m in range(M)]).astype(np.int64)
pix = np.concatenate([score[m, :count[m]] for m in range(M)]).astype(np.int32)
n_pairs = len(pair_mask)

# host-side inputs of the shape score: targets + gradient images in pinned memory
t_arr, t_ptr = ctx.host_alloc(T * img)
g_arr, g_ptr = ctx.host_alloc(T * 2 * W * H)
targets = t_arr.reshape(T, H, W, 3)
grads = g_arr.view(np.uint16).reshape(T, H, W)
for i in range(0, T, 64):
    n = min(64, T - i)
    targets[i:i + n] = ctx.synth_rgb(1, SEED, i, n, W, H, on_device=True)
    grads[i:i + n] = ctx.synth_gradient(SEED, i, n, W, H, on_device=True)

t0 = time.perf_counter()
sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
sms.add_rgb(masks)
prep_s = time.perf_counter() - t0
sms.score_pairs(targets[:64], grads[:64], None, pair_mask[:8] * 0, pair_target[:8] % 64)      # warm-up (buffers)
t0 = time.perf_counter()
gap, he, mir = sms.score_pairs(targets, grads, None, pair_mask, pair_target)
e2e_s = time.perf_counter() - t0
st = ctx.last_stats()
norm_t0 = time.perf_counter()
normalized = np.concatenate([capi.normalize_scores(pix[pair_mask == m], gap[pair_mask == m], he[pair_mask == m]) for m in range(min(M, 50))])
norm_s = (time.perf_counter() - norm_t0) / max(1, min(M, 50)) * M

rng = np.random.default_rng(3)
ok = True
for i in rng.choice(n_pairs, min(a.check_pairs, n_pairs), replace=False):
    m, t = int(pair_mask[i]), int(pair_target[i])
    om = O.ShapeMask(masks[m], 20, True, rects)
    exp = om.score(targets[t], grads[t], O.make_zgap(targets[t], 20, rects))
    ok &= (int(gap[i]), int(he[i]), bool(mir[i])) == exp
print(json.dumps({
    "workload": "BASELINE configs[2]: shape score of the top-%d pixel matches of each of %d masks over %d targets" % (K, M, T),
    "pairs": int(n_pairs), "distinct_targets": int(len(np.unique(pair_target))),
    "pairs_per_s_e2e": n_pairs / e2e_s, "e2e_s": e2e_s, "pair_kernel_ms": st["match_kernel_ms"],
    "pairs_per_s_kernel": n_pairs / (st["match_kernel_ms"] * 1e-3), "h2d_bytes": int(st["h2d_bytes"]),
    "mask_prep_s": prep_s, "mask_prep_ms_per_mask": prep_s / M * 1e3, "normalize_s_estimate": norm_s,
    "parity_sampled_pairs": bool(ok), "checked_pairs": int(min(a.check_pairs, n_pairs))}))
sms.close()
