"""The single-pair provider alone (bench.py's pair_provider leg) for several caller-thread counts:  python tools/pairq_bench.py [threads ...]"""
import json, sys
import numpy as np
sys.path.insert(0, ".")
import bench
from colormipsearch_b200 import capi

threads = [int(x) for x in sys.argv[1:]] or [40, 80]
ctx = capi.Context(device_ids=[0])
masks = np.concatenate([ctx.synth_rgb(0, bench.SEED, i, 64, bench.W, bench.H, on_device=True) for i in range(0, 64, 64)])
for t in threads:
    r = bench.pair_provider_bench(ctx, masks, threads=t)
    print(json.dumps({"threads": t, "warm": r["warm"]["value"], "cold": r["cold"]["value"], "mean_batch": r["warm"]["mean_batch"], "ok": r["equals_dense_search"]}), flush=True)
