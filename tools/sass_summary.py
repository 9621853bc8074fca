"""SASS evidence per kernel of libsblx.so (runs without a GPU): registers / shared / local memory from
`cuobjdump -res-usage`, instruction counts from `cuobjdump -sass` (256-bit global accesses, FP64 math, shared-memory
loads, shuffles, barriers, local-memory spill traffic, bulk L2 prefetch).
  python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sparlectra.jl_b200", "lib", "libsblx.so")
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
name = None
for ln in res.splitlines():
    m = re.search(r"Function (\S+):", ln)
    if m:
        name = m.group(1)
    m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", ln)
    if m and name:
        usage[name] = tuple(int(x) for x in m.groups())
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
PAT = [("LDG.256", r"\bLDG\.E[.\w]*\.256"), ("LDG.128", r"\bLDG\.E[.\w]*\.128"), ("LDG.64", r"\bLDG\.E[.\w]*\.64\b"), ("LDG.other", r"\bLDG\.E(?![.\w]*\.(256|128|64)\b)"),
       ("STG.256", r"\bSTG\.E[.\w]*\.256"), ("STG.128", r"\bSTG\.E[.\w]*\.128"), ("STG.other", r"\bSTG\.E(?![.\w]*\.(256|128)\b)"),
       ("DFMA", r"\bDFMA\b"), ("DMUL", r"\bDMUL\b"), ("DADD", r"\bDADD\b"), ("MUFU.RCP64H", r"MUFU\.RCP64H"),
       ("LDS.128", r"\bLDS[.\w]*\.128"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("SHFL", r"\bSHFL\b"), ("BAR", r"\bBAR\b"),
       ("LDGSTS", r"\bLDGSTS\b"), ("UBLKPF", r"\bUBLKPF\b"), ("LDL(spill)", r"\bLDL\b"), ("STL(spill)", r"\bSTL\b"), ("ATOM/RED", r"\b(ATOMG|RED|ATOM)\b")]
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur and "/*" in ln:
        counts[cur]["inst"] += 1 if re.search(r"/\*[0-9a-f]{4,}\*/", ln) else 0
        for key, pat in PAT:
            if re.search(pat, ln):
                counts[cur][key] += 1
def dem(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void sblx::", "").replace("sblx::", "")
    except Exception:
        return n
print("# libsblx.so SASS summary (sm_100a, static instruction counts per kernel; nvcc " +
      subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2].split(",")[1].strip() + ")")
cols = ["inst"] + [k for k, _ in PAT]
print("%-28s %4s %6s %6s %5s | " % ("kernel", "regs", "stack", "shared", "local") + " ".join("%9s" % c for c in cols))
for fn, c in counts.items():
    u = usage.get(fn, (0, 0, 0, 0))
    print("%-28s %4d %6d %6d %5d | " % (dem(fn)[:28], u[0], u[1], u[2], u[3]) + " ".join("%9d" % c[k] for k in cols))
print("\nLDG.256 / STG.256 = one 2x2 FP64 block (32 B, one DRAM sector) per instruction; LDL/STL = local-memory (spill) traffic;")
print("'stack' bytes are the per-thread frame (24 B frames are saved predicates of the pivot chain, no loop-carried spills).")
