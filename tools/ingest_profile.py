"""A fixed workload for profiling the ingest kernels: 2048 synthetic targets as PackBits TIFF files, streamed against 1000 masks
(cds_search_stream_tiff), twice.  Prints the wall time of the second call."""
import sys, time
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, ".")
from colormipsearch_b200 import capi
from oracle import oracle as O

W, H, SEED = 1210, 566, 0xC0FFEE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rects = O.label_rects(W, H)
ctx = capi.Context(device_ids=[0])
ctx.set_option("fused_ingest", fused)
if len(sys.argv) > 3:
    ctx.set_option("stream_chunk_tiff", int(sys.argv[3]))
masks = np.concatenate([ctx.synth_rgb(0, SEED, i, min(64, 1000 - i), W, H, on_device=True) for i in range(0, 1000, 64)])
targets = np.concatenate([ctx.synth_rgb(1, SEED, i, 64, W, H, on_device=True) for i in range(0, n, 64)])
with ThreadPoolExecutor(16) as ex:
    files = list(ex.map(lambda t: capi.tiff_encode_rgb(t, 8, 32773), targets))
off = np.zeros(n + 1, np.int64)
np.cumsum([len(f) for f in files], out=off[1:])
arr, ptr = ctx.host_alloc(int(off[-1]) + 64)
for i, f in enumerate(files):
    arr[off[i]:off[i + 1]] = np.frombuffer(f, np.uint8)
ms = capi.MaskSet(ctx, W, H, 20, 20, 0.01, 2, True, rects)
ms.add_rgb(masks)
for _ in range(2):
    t0 = time.perf_counter()
    r = ms.search_stream_tiff((arr, off), 300, 1.0, blob_ptr=ptr)
    dt = time.perf_counter() - t0
print("fused", fused, "targets", n, "ms", dt * 1e3, "device ms", ctx.last_stats()["total_device_ms"], "match ms", ctx.last_stats()["match_kernel_ms"])
