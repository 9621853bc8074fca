"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys, collections, re
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None
per = collections.OrderedDict()
tot = toti = 0
for r in rows:
    if not r: continue
    if r[0] in ("Kernel Name", "File Name", "Function Name"): continue
    if "Source" in r and ("Warp Stall Sampling (All Samples)" in r or "# Samples" in r):
        hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    src = d.get("Source", "")
    first = r[0]
    # cuda lines have a numeric line number in first column and no address
    if re.fullmatch(r"\d+", first.strip() or "x") and not first.startswith("0x"):
        cur = (int(first), src.strip()); per.setdefault(cur, [0, 0]); 
        try:
            s = int(d.get("Warp Stall Sampling (All Samples)", "0") or 0); i = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            s = i = 0
        per[cur][0] += s; per[cur][1] += i; tot += s; toti += i
print("total samples", tot, "inst", toti)
for (ln, src), (s, i) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  L%-4d %s" % (100 * s / max(tot, 1), 100 * i / max(toti, 1), ln, src[:120]))
