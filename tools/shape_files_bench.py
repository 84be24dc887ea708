"""The shape score end to end with its inputs as pixels, as TIFF files + gradient pixels, and as TIFF + PNG files (gradient streams
inflated on the device or by host threads), on a synthetic pair mix of ~8 pairs per target: bench.py's shape_config2_mix without the
pixel-match search in front of it.

    python tools/shape_files_bench.py [--masks 64] [--targets 2048] [--per-mask 256]
"""
import argparse, json, sys
import numpy as np
sys.path.insert(0, ".")
import bench
from colormipsearch_b200 import capi

ap = argparse.ArgumentParser()
ap.add_argument("--masks", type=int, default=64)
ap.add_argument("--targets", type=int, default=4096)
ap.add_argument("--per-mask", type=int, default=512)
ap.add_argument("--windows", default="0,1024,512", help="cds_ctx_set_option shape_inflate_window settings to time (0 = default)")
a = ap.parse_args()
rng = np.random.default_rng(11)
target = np.stack([rng.choice(a.targets, a.per_mask, replace=False) for _ in range(a.masks)]).astype(np.int64)
count = np.full(a.masks, a.per_mask, np.int64)
with capi.Context(device_ids=[0]) as ctx:
    out = bench.shape_config2_mix(ctx, (None, target, None, count), a.masks, 0, a.targets, tuple(int(x) for x in a.windows.split(",")))
print(json.dumps(out))
