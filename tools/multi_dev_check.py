"""One cds_ctx over all visible devices: dense / top-K / streaming searches (pixels and TIFF files, library from TIFF files, all
matches) must equal the single-device results."""
import sys, numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O
W, H = 1210, 566
masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 24, W, H)
targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 200, W, H)
rects = O.label_rects(W, H)
files = [capi.tiff_encode_rgb(t, 8 if i % 3 else 566, 32773 if i % 5 else 1) for i, t in enumerate(targets)]
res = {}
for nd in (1, 0):
    ctx = capi.Context(n_dev=nd)
    print("devices", ctx.num_devices)
    ctx.set_option("stream_chunk", 16)
    ctx.set_option("stream_chunk_tiff", 24)
    lib = capi.Library(ctx, W, H, 256); lib.add_rgb(targets)
    ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects); ms.add_rgb(masks)
    lib2 = capi.Library(ctx, W, H, 256); lib2.add_tiff(files)
    res[nd] = (ms.search_dense(lib), ms.search_topk(lib, 50, 0.0), ms.search_stream(targets, 50, 0.0), ms.search_stream_tiff(files, 50, 0.0),
               ms.search_topk(lib2, 50, 0.0), ms.search_stream_matches_tiff(files, 1.0), ms.search_matches(lib, 1.0))
    ms.close(); lib.close(); lib2.close(); ctx.close()
a, b = res[1], res[0]
print("dense equal:", np.array_equal(a[0][0], b[0][0]), np.array_equal(a[0][1], b[0][1]))
for name, i in (("topk", 1), ("stream", 2), ("stream_tiff", 3), ("topk over a TIFF-built library", 4)):
    ok = np.array_equal(a[i][3], b[i][3]) and all(np.array_equal(a[i][j][m, :a[i][3][m]], b[i][j][m, :a[i][3][m]]) for j in range(3) for m in range(24))
    print(name, "equal:", ok)
ok = all(np.array_equal(a[1][j][m, :a[1][3][m]], a[2][j][m, :a[1][3][m]]) for j in range(3) for m in range(24))
print("stream == topk:", ok)
for i in (3, 4):
    ok = all(np.array_equal(a[1][j][m, :a[1][3][m]], b[i][j][m, :a[1][3][m]]) for j in range(3) for m in range(24))
    print("multi-device result %d == single-device topk:" % i, ok)
print("all-matches over TIFF files == all-matches over the resident library (single, multi):",
      all(np.array_equal(x, y) for x, y in zip(a[5], a[6])), all(np.array_equal(x, y) for x, y in zip(b[5], b[6])),
      "; single == multi:", all(np.array_equal(x, y) for x, y in zip(a[5], b[5])))
