"""Rank CUDA source lines of an ncu report by executed instructions / stall samples:  python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr = rows[hi]
ii = hdr.index("Instructions Executed"); si = hdr.index("# Samples")
def f(x):
    try: return float(x)
    except Exception: return 0.0
lines = [r for r in rows[hi + 1:] if r and r[0] != "" and len(r) > max(ii, si) and r[0].isdigit()]
tot = sum(f(r[ii]) for r in lines); tots = sum(f(r[si]) for r in lines)
print("total inst %.4g samples %d" % (tot, tots))
for r in lines:
    v = f(r[ii]); s = f(r[si])
    if v / tot * 100 > thr or s / tots * 100 > thr:
        print("%4s %5.1f%% inst %5.1f%% samp | %s" % (r[0], v / tot * 100, s / tots * 100, r[1].strip()[:120]))
