"""Candidate kernel: executed instructions and stall samples of an ncu report by PHASE of the kernel (source line ranges of cds_cand.cu),
every SASS instruction counted once:  python tools/ncu_phases.py report.ncu-rep"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [r for r in rows if r and r[0] == "Line No"][0]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
seen = {}
cur_file = cur_line = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == "" and len(r) > 40 and r[2].startswith("0x"):
        seen.setdefault(r[2], {"att": [], "row": r})["att"].append((cur_file, cur_line))
def num(x):
    try: return float(x)
    except Exception: return 0.0
RANGES = [((79, 92), "evaluation"), ((118, 137), "evaluation"), ((163, 183), "evaluation"), ((139, 152), "expansion: select_bit"), ((154, 161), "palette reference load")]
OUTER = [((292, 297), "wait for band (call site)"), ((319, 319), "wait for band (call site)"), ((334, 342), "submit"), ((345, 396), "expansion: peel"),
         ((399, 407), "submit_words"), ((425, 446), "scan of passing tickets"), ((447, 485), "ticket test"), ((486, 504), "scan of passing tickets"),
         ((505, 526), "band flush + release"), ((527, 545), "item epilogue"), ((296, 333), "band setup"), ((408, 424), "band setup"), ((235, 280), "producer")]
def phase(att):
    files = dict(att)
    if files.get("cds_ptx.cuh") in range(25, 60): return "wait for band (try_wait loop)"
    ls = [l for f, l in att if f == "cds_cand.cu"]
    if not ls: return "intrinsics / other headers"
    for (a, b), name in RANGES:
        if any(a <= x <= b for x in ls): return name
    l = min(ls)
    for (a, b), name in OUTER:
        if a <= l <= b: return name
    return "kernel prologue / misc"
inst = collections.Counter(); samp = collections.Counter(); st = collections.defaultdict(collections.Counter)
for a, v in seen.items():
    r = v["row"]; ph = phase(v["att"])
    inst[ph] += num(r[col["Instructions Executed"]]); samp[ph] += num(r[col["# Samples"]])
    for s in stalls: st[ph][s] += num(r[col[s]])
ti = sum(inst.values()); ts = sum(samp.values())
print("%-34s %7s %8s   top stall reasons (share of the phase's samples)" % ("phase", "inst %", "samples %"))
for ph, v in samp.most_common():
    top = ", ".join("%s %.0f%%" % (k[6:], c / max(1.0, sum(st[ph].values())) * 100) for k, c in st[ph].most_common(4))
    print("%-34s %6.1f%% %7.1f%%   %s" % (ph, inst[ph] / ti * 100, v / ts * 100, top))
