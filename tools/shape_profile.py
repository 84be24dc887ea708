"""A small, fixed workload for profiling the kernels next to the candidate kernel (ncu launch lists / --set full captures):
shape mask preparation of 64 masks, shape scoring of 64 targets (zgap derived on the device) with 8 pairs per target, and one
streamed pixel-match search over 512 host targets (encoder, occupancy, top-K kernels)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O

W, H, SEED = 1210, 566, 0xC0FFEE
rects = O.label_rects(W, H)
ctx = capi.Context(device_ids=[0])
masks = ctx.synth_rgb(0, SEED, 0, 64, W, H, on_device=True)
targets = np.concatenate([ctx.synth_rgb(1, SEED, i, 64, W, H, on_device=True) for i in range(0, 512, 64)])
grads = ctx.synth_gradient(SEED, 0, 64, W, H, on_device=True)
sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
sms.add_rgb(masks)
pm = np.repeat(np.arange(64, dtype=np.int32), 8)
pt = (np.arange(512, dtype=np.int64) * 7) % 64
for _ in range(2):
    gap, he, mir = sms.score_pairs(targets[:64], grads, None, pm, pt)
sms.close()
ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
ms.add_rgb(masks)
for _ in range(2):
    ms.search_stream(targets, 300, 1.0)
ms.close()
print("ok", int(gap.sum()))
