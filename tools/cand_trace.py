"""Reads a CDSGPU_CAND_TRACE dump (SM clocks of CTA 0's first items: per band, when every consumer warp got the band and when it
released it, and when the producer could issue the band's loads) and prints where a band's time goes.

    CDSGPU_CAND_TRACE=gpurun_out/trace.bin python tools/cand_sweep.py --targets 256 --reps 1 --settings wait=0,hint=0
    python tools/cand_trace.py gpurun_out/trace.bin
"""
import sys
import numpy as np
raw = open(sys.argv[1], "rb").read()
items, maxb, nb, ncw = np.frombuffer(raw[:16], np.int32)
t = np.frombuffer(raw[16:], np.int64).reshape(items, maxb, 32, 2)[:, :nb]
start, done = t[:, :, :ncw, 0], t[:, :, :ncw, 1]
issue = t[:, :, 31, 0]
ok = (start > 0).all(axis=2) & (done > 0).all(axis=2)
rows = []
for i in range(1, items):            # item 0 includes start-up
    for b in range(nb):
        if not ok[i, b]:
            continue
        s, d = start[i, b], done[i, b]
        rows.append((d.max() - s.min(), (d - s).mean(), (d - s).min(), (d - s).max(), s.max() - s.min(), d.max() - d.min(),
                     s.min() - issue[i, b] if issue[i, b] > 0 else 0))
r = np.array(rows, float)
names = ["band span (first start -> last done)", "mean busy per warp", "min busy", "max busy", "spread of starts", "spread of dones", "load issue -> first start"]
print("bands analysed:", len(r), " item span (clocks):", float((done[1:, nb - 1].max(axis=1) - start[1:, 0].min(axis=1)).mean()))
for n, col in zip(names, r.T):
    print("%-40s mean %9.0f  p50 %9.0f  p90 %9.0f  max %9.0f" % (n, col.mean(), np.median(col), np.percentile(col, 90), col.max()))
busy = (done - start)[1:][ok[1:]].sum()
span = 0.0
for i in range(1, items):
    span += (done[i, nb - 1].max() - start[i, 0].min()) * ncw
print("busy fraction of consumer-warp time inside items: %.3f" % (busy / span))
