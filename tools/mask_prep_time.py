import sys, time, numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O
W, H = 1210, 566
M = 1000
ctx = capi.Context(device_ids=[0])
rects = O.label_rects(W, H)
img = 3 * W * H
arr, ptr = ctx.host_alloc(M * img)
for i in range(0, M, 64):
    n = min(64, M - i)
    arr[i * img:(i + n) * img] = ctx.synth_rgb(0, 0xC0FFEE, i, n, W, H, on_device=True).reshape(-1)
for rep in range(4):
    t0 = time.perf_counter()
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
    ms.add_rgb_ptr(ptr, M)
    t1 = time.perf_counter()
    ms.close()
    print("add_rgb %d masks: %.1f ms" % (M, (t1 - t0) * 1e3))
