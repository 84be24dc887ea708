"""Where a step's mask work goes: create + add (pixels or PackBits TIFF files in pinned host memory), then the first search, which builds
the palettes and the candidate kernel's word lists (cds_maskset::sync_descs).   python tools/mask_prep_time.py [n_masks]"""
import sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O
W, H = 1210, 566
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = capi.Context(device_ids=[0])
rects = O.label_rects(W, H)
img = 3 * W * H
arr, ptr = ctx.host_alloc(M * img)
masks = np.concatenate([ctx.synth_rgb(0, 0xC0FFEE, i, min(64, M - i), W, H, on_device=True) for i in range(0, M, 64)])
arr[:] = masks.reshape(-1)
with ThreadPoolExecutor(16) as ex:
    files = list(ex.map(lambda t: capi.tiff_encode_rgb(t, 8, 32773), masks))
off = np.zeros(M + 1, np.int64)
np.cumsum([len(f) for f in files], out=off[1:])
farr, fptr = ctx.host_alloc(int(off[-1]) + 64)
for i, f in enumerate(files):
    farr[off[i]:off[i + 1]] = np.frombuffer(f, np.uint8)
lib = capi.Library(ctx, W, H, 64)
lib.generate_synthetic(0xC0FFEE, 0, 64)
for rep in range(4):
    for kind in ("pixels", "tiff"):
        t0 = time.perf_counter()
        ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
        t1 = time.perf_counter()
        if kind == "pixels":
            ms.add_rgb_ptr(ptr, M)
        else:
            ms.add_tiff((farr, off), blob_ptr=fptr)
        t2 = time.perf_counter()
        ms.search_topk(lib, 8, 1.0)
        t3 = time.perf_counter()
        ms.search_topk(lib, 8, 1.0)
        t4 = time.perf_counter()
        ms.close()
        t5 = time.perf_counter()
        print("%-6s create %.2f  add %.2f  first search (64 targets) %.2f  second %.2f  -> list build %.2f  close %.2f ms" %
              (kind, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t3 - t2 - (t4 - t3)) * 1e3, (t5 - t4) * 1e3), flush=True)
