"""One tick isolated from an ncu per-launch metrics CSV (tools/r2_call7.sh) -> per-kernel time / DRAM bytes / share and the
k_front DRAM bytes per scenario-iteration that bench.py reads for `roofline.traffic`.
  python tools/traffic_json.py gpurun_out/r2_tick_metrics.csv 2048 > profiles/r2_k_front_traffic.json"""
import collections, csv, json, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
units = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rows = collections.OrderedDict()
for row in csv.DictReader(lines):
    d = rows.setdefault(int(row["ID"]), {"name": re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")})
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; m = row["Metric Name"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    if "bytes" in m:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    d[m] = v
ids = sorted(rows)
starts = [i for i in ids if rows[i]["name"].startswith("k_assign")]
tick = [i for i in ids if starts[-1] <= i] if len(starts) == 1 else [i for i in ids if starts[-2] <= i < starts[-1]]
ker = collections.OrderedDict()
for i in tick:
    d = rows[i]
    base = "k_front" if d["name"].startswith("k_front") else "k_backsolve" if d["name"].startswith("k_backsolve<") else d["name"]
    k = ker.setdefault(base, {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "fp64_pct_x_us": 0.0, "l2_hit_x_us": 0.0})
    us = d["gpu__time_duration.sum"]
    k["launches"] += 1; k["us"] += us
    k["dram_bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    k["fp64_pct_x_us"] += us * d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0)
    k["l2_hit_x_us"] += us * d.get("lts__t_sector_hit_rate.pct", 0)
T = sum(k["us"] for k in ker.values())
for k in ker.values():
    k["share"] = round(k["us"] / T, 4)
    k["dram_bytes_per_unit"] = round(k["dram_bytes"] / units, 1)
    k["dram_GBps"] = round(k["dram_bytes"] / max(k["us"], 1e-9) / 1e3, 1)
    k["fp64_pipe_pct"] = round(k.pop("fp64_pct_x_us") / max(k["us"], 1e-9), 1)
    k["l2_hit_pct"] = round(k.pop("l2_hit_x_us") / max(k["us"], 1e-9), 1)
    k["us"] = round(k["us"], 1)
out = {"command": "python bench.py --scenarios %d --steps 1 --warmup 0 --no-cpu-baseline --no-api-e2e (one tick of the batch, SBLX_PROFILE_TICK=3; "
                  "ncu --metrics ... --clock-control none; launch times are serialised / cold-cache: compare shares)" % units,
       "build": "round-2 (ABI 0.2.0: same factor / solve kernels as round 1 plus the tiny-pivot guard)",
       "units_in_tick": units, "tick_us": round(T, 1), "kernels": ker,
       "k_front_dram_bytes_per_unit": ker["k_front"]["dram_bytes_per_unit"], "k_front_share": ker["k_front"]["share"]}
print(json.dumps(out, indent=1))
