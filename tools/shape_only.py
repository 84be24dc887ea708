import sys, json, time
sys.path.insert(0, ".")
import bench
from colormipsearch_b200 import capi
ctx = capi.Context(device_ids=[0])
for _ in range(2):
    out = bench.shape_bench(ctx, cpu_pairs=0) if False else bench.shape_bench(ctx)
print(json.dumps({k: out[k] for k in ("value", "kernel_ms", "e2e", "mask_prep_ms_per_mask")}))
