"""Warp-stall samples of k_front in an .ncu-rep grouped by kernel PHASE (source line ranges between the PHASE_MARK
lines of kernels.cuh) and by stall reason.  usage: python tools/ncu_phases.py report.ncu-rep"""
import csv, subprocess, sys, collections, re, os
rep = sys.argv[1]
src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sparlectra.jl_b200", "csrc", "kernels.cuh")).read().split("\n")
marks = [(i + 1, re.search(r"PHASE_MARK\((\d)\)", l).group(1)) for i, l in enumerate(src) if re.search(r"^\s*PHASE_MARK\(\d\)", l)]
kstart = next(i + 1 for i, l in enumerate(src) if "k_front(DevModel" in l)
names = {"0": "ingest pivot blk", "1": "chain (warp0) + ingest rest", "2": "barrier after chain", "3": "substitution", "4": "U store",
         "5": "Schur", "6": "update vector", "7": "tail"}
def phase_of(line):
    if line < kstart: return "helpers(inlined: cadd/bmsub/...)"
    for ln, ph in marks:
        if line <= ln: return names[ph]
    return "after"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
hdr = None
agg = collections.OrderedDict()
reasons = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_math", "stall_no_inst", "stall_selected", "stall_not_selected", "stall_lg", "stall_mio"]
for r in csv.reader(out.splitlines()):
    if not r: continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not re.fullmatch(r"\d+", r[0].strip()): continue
    d = dict(zip(hdr, r))
    ph = phase_of(int(r[0]))
    a = agg.setdefault(ph, collections.Counter())
    try:
        a["all"] += int(d.get("Warp Stall Sampling (All Samples)") or 0)
        a["inst"] += int(d.get("Instructions Executed") or 0)
        for k in reasons: a[k] += int(d.get(k) or 0)
    except ValueError:
        pass
tot = sum(a["all"] for a in agg.values()); toti = sum(a["inst"] for a in agg.values())
print("%-36s %7s %7s | " % ("phase", "smp%", "inst%") + " ".join("%9s" % k.replace("stall_", "") for k in reasons))
for ph, a in agg.items():
    print("%-36s %6.1f%% %6.1f%% | " % (ph, 100 * a["all"] / max(tot, 1), 100 * a["inst"] / max(toti, 1)) + " ".join("%8.1f%%" % (100 * a[k] / max(tot, 1)) for k in reasons))
