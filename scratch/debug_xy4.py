import sys, numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O
W, H = 1210, 566
ctx = capi.Context(n_dev=1)
masks = capi.synth_rgb_host(0, 0xC0FFEE, 0, 24, W, H)
targets = capi.synth_rgb_host(1, 0xC0FFEE, 0, 70, W, H)
lib = capi.Library(ctx, W, H, 80); lib.add_rgb(targets)
rects = O.label_rects(W, H)
for (mthr, dthr, ztol, xys, mirror) in [(20, 20, 0.005, 4, False), (20, 20, 0.005, 4, True), (20, 20, 0.01, 2, True)]:
    oms = [O.PixelMatchMask(x, mthr, mirror, dthr, ztol, xys, rects) for x in masks]
    es, em, _ = O.search_dense(oms, targets)
    for kern in ("cand", "band"):
        ctx.set_match_kernel(kern)
        ms = capi.MaskSet(ctx, W, H, mthr, dthr, ztol, xys, mirror, rects)
        ms.add_rgb(masks)
        s, m = ms.search_dense(lib)
        st = ctx.last_stats()
        bad = np.argwhere(s != es)
        print(xys, mirror, kern, "kernel", st["match_kernel"], "mismatches", len(bad), "of", s.size, "sum got", int(s.sum()), "exp", int(es.sum()))
        for (a, b) in bad[:6]:
            vs = oms[a].variant_scores(targets[b])
            print("   m", a, "t", b, "got", s[a, b], "exp", es[a, b], "variants", np.asarray(vs).tolist())
        ms.close()
