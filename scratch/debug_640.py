import numpy as np, sys, os
sys.path.insert(0, '/root/repo')
from colormipsearch_b200 import capi
from oracle import oracle as O
w, h = 640, 480
rng = np.random.default_rng(w * 1000 + h)
lut = O.lut().astype(np.uint8)
def rand_img(density):
    img = np.zeros((h, w, 3), np.uint8)
    sel = rng.random((h, w)) < density
    z = rng.integers(0, 256, (h, w))
    br = rng.integers(20, 256, (h, w, 1))
    col = (lut[z].astype(np.int64) * br // 255).astype(np.uint8)
    img[sel] = col[sel]
    return img
masks = np.stack([rand_img(0.2) for _ in range(18)])
targets = np.stack([rand_img(0.5) for _ in range(9)])
rects = np.array([[0, 0, w // 3, h // 4]], np.int32)
ctx = capi.Context(n_dev=1)
lib = capi.Library(ctx, w, h, 9); lib.add_rgb(targets)
params = (20, 20, 0.02, 2, True)
mthr, dthr, ztol, xys, mirror = params
oms = [O.PixelMatchMask(x, mthr, mirror, dthr, ztol, xys, rects) for x in masks]
es, em, _ = O.search_dense(oms, targets)
print("P:", [m.size for m in oms])
ms = capi.MaskSet(ctx, w, h, mthr, dthr, ztol, xys, mirror, rects); ms.add_rgb(masks)
sb, mb = ms.search_dense(lib)
print("band mismatches:", np.argwhere(sb != es).tolist(), (sb-es)[sb!=es].tolist())
sg = np.zeros_like(sb)
for i in range(0, 18, 2):
    m2 = capi.MaskSet(ctx, w, h, mthr, dthr, ztol, xys, mirror, rects); m2.add_rgb(masks[i:i+2])
    s2, _ = m2.search_dense(lib); sg[i:i+2] = s2; m2.close()
print("gather mismatches:", np.argwhere(sg != es).tolist(), (sg-es)[sg!=es].tolist())
for (m, t) in np.argwhere(sb != es).tolist()[:3]:
    print("cell", m, t, "oracle variants", oms[m].variant_scores(targets[t]).tolist(), "band", sb[m,t], "gather", sg[m,t])
