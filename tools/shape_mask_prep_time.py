import sys, time
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O
W, H = 1210, 566
M = 1000
ctx = capi.Context(device_ids=[0])
rects = O.label_rects(W, H)
arr, ptr = ctx.host_alloc(M * 3 * W * H)
for i in range(0, M, 64):
    n = min(64, M - i)
    arr[i * 3 * W * H:(i + n) * 3 * W * H] = ctx.synth_rgb(0, 0xC0FFEE, i, n, W, H, on_device=True).reshape(-1)
for rep in range(5):
    t0 = time.perf_counter()
    sms = capi.ShapeMaskSet(ctx, W, H, 20, True, rects)
    n = 64 if rep == 0 else M
    sms.add_rgb_ptr(ptr, n)
    t1 = time.perf_counter()
    sms.close()
    t2 = time.perf_counter()
    print("rep", rep, "masks", n, "add %.1f ms (%.3f ms/mask)  close %.1f ms" % ((t1 - t0) * 1e3, (t1 - t0) * 1e3 / n, (t2 - t1) * 1e3), flush=True)
