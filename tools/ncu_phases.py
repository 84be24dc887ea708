"""Candidate kernel: executed instructions and stall samples of an ncu report by PHASE of the kernel (source line ranges of cds_cand.cu),
every SASS instruction counted once:  python tools/ncu_phases.py report.ncu-rep"""
import collections, csv, os, subprocess, sys
# Phases by MARKER lines of cds_cand.cu (the first line that contains the text starts the phase; it lasts until the next marker), so
# that the table survives edits of the kernel.  INNER phases are device functions: an instruction inlined from one of them counts there.
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "colormipsearch_b200", "csrc", "cds_cand.cu")
INNER_MARKS = [("__device__ __forceinline__ void count_hit", "evaluation"), ("struct CandSmem", None), ("struct EvalUnroll", "evaluation"),
               ("__device__ __forceinline__ uint32_t select_bit", "expansion: select_bit"), ("__device__ __forceinline__ uint32_t fetch_palette_ref", "palette reference load"),
               ("__device__ __forceinline__ void eval_candidates", "evaluation"), ("__global__ void __launch_bounds__((NCW + 1) * 32, 1) pixelmatch_cand_kernel", None)]
OUTER_MARKS = [("pixelmatch_cand_kernel(const CandParams p)", "kernel prologue / misc"), ("if (warp == NCW) {", "producer"), ("// ---------------------------------------------------------------------- consumers", "kernel prologue / misc"),
               ("auto wait_full = ", "wait for band (call site)"), ("    for (;;) {\n        wait_full();", "item / band setup"), ("auto run_pending = ", "submit"),
               ("auto peel = ", "expansion: peel"), ("auto drain_words = ", "drain of the word queue"), ("// Tickets.  Inside a tile row", "item / band setup"),
               ("auto scan_one = ", "scan of passing tickets"), ("uint32_t bt = 0;", "ticket test"), ("// scan the tickets that passed", "scan of passing tickets"),
               ("// the band's last, partly filled batches", "band flush + release"), ("// item epilogue", "item epilogue"), ("struct CandConfig", None)]
def mark_lines(marks):
    text = open(SRC).read()
    out = []
    for m, name in marks:
        k = text.find(m)
        if k < 0: raise SystemExit("marker not found in cds_cand.cu: " + m)
        out.append((text.count("\n", 0, k) + 1, name))
    return sorted(out)
INNER, OUTER = mark_lines(INNER_MARKS), mark_lines(OUTER_MARKS)
def lookup(table, line):
    name = None
    for l, n in table:
        if l <= line: name = n
        else: break
    return name
def phase(att):
    files = dict(att)
    if files.get("cds_ptx.cuh") in range(25, 60): return "wait for band (try_wait loop)"
    ls = [l for f, l in att if f == "cds_cand.cu"]
    if not ls: return "intrinsics / other headers"
    for x in ls:
        n = lookup(INNER, x)
        if n and x < OUTER[0][0]: return n
    return lookup(OUTER, max(ls)) or "kernel prologue / misc"

def main():
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


if __name__ == "__main__":
    main()
