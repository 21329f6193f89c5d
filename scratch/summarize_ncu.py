"""turns gpurun_out/*.ncu-rep and the launch list into the small tracked summaries under profiles/"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]

def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

def report(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = to_bytes(r[i], units[i]) if "bytes" in k else float(r[i].replace(",", "") or 0)
                if k == "gpu__time_duration.sum":
                    d[k + "_unit"] = units[i]
        d["dram_traffic_bytes"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        out.append(d)
    json.dump(out, open(os.path.join(OUT, f"{tag}.json"), "w"), indent=1)
    return out

def launches(csvpath, tag):
    rows = list(csv.reader(l for l in open(csvpath) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        if len(r) > vi:
            try: agg[r[ki]].append(float(r[vi].replace(",", "")))
            except ValueError: pass
    total = sum(sum(v) for v in agg.values())
    lines = ["kernel,launches,avg_ns,min_ns,total_ns,share_of_all_launch_time"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"\"{k[:110]}\",{len(v)},{sum(v)/len(v):.0f},{min(v):.0f},{sum(v):.0f},{sum(v)/total:.4f}")
    open(os.path.join(OUT, f"{tag}.csv"), "w").write("\n".join(lines) + "\n")
    return agg

if __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    for rep, tag in [("prof_c2_r1.ncu-rep", "r1_ncu_full_c2_per_kernels"), ("prof_c3_r1.ncu-rep", "r1_ncu_full_c3_per_kernels"),
                     ("prof_c4_r1.ncu-rep", "r1_ncu_full_c4_scan_kernels"),
                     ("prof_c4_r1b.ncu-rep", "r1_ncu_full_c4_tma_scan_kernels"),
                     ("prof_c2_r1b.ncu-rep", "r1_ncu_full_c2_per_kernels_pdl"), ("prof_c3_r1b.ncu-rep", "r1_ncu_full_c3_per_kernels_pdl"), ("prof_scan_r1e.ncu-rep", "r1_ncu_full_wtile_before_fastpath"),
                     ("prof_c2_s7.ncu-rep", "r1_ncu_full_c2_per_kernels_s7"), ("prof_c3_s7.ncu-rep", "r1_ncu_full_c3_per_kernels_s7"),
                     ("prof_c4_s7.ncu-rep", "r1_ncu_full_c4_tma_scan_kernels_s7")]:
        p = os.path.join(g, rep)
        if os.path.exists(p):
            for d in report(p, tag):
                print(tag, d["kernel"][:60], d["gpu__time_duration.sum"], d.get("gpu__time_duration.sum_unit"), round(d["dram_traffic_bytes"] / 1e6, 2), "MB")
    # per-launch traffic of the TMA scan kernels -> profiles/r1_traffic.json (bench.py's roofline.traffic)
    tp = os.path.join(OUT, "r1_traffic.json")
    tj = os.path.join(OUT, "r1_ncu_full_c4_tma_scan_kernels_s7.json")
    if os.path.exists(tp) and os.path.exists(tj):
        t = json.load(open(tp))
        for d in json.load(open(tj)):
            name = d["kernel"].split("::")[-1].split("(")[0]
            t["c4:" + name] = {"ncu_duration_us": round(d["gpu__time_duration.sum"], 2),
                               "dram_traffic_bytes": int(d["dram_traffic_bytes"]),
                               "dram_read_bytes": int(d["dram__bytes_read.sum"]),
                               "dram_write_bytes": int(d["dram__bytes_write.sum"]), "launches_captured": 1,
                               "source": "profiles/r1_ncu_full_c4_tma_scan_kernels_s7.json (ncu --set full --clock-control none; cold caches, serialised)"}
        json.dump(t, open(tp, "w"), indent=1)
    for cfg, tag in (("c2", "r1_ncu_full_c2_per_kernels_s7"), ("c3", "r1_ncu_full_c3_per_kernels_s7")):
        tj = os.path.join(OUT, tag + ".json")
        if os.path.exists(tp) and os.path.exists(tj):
            t = json.load(open(tp))
            seen = collections.defaultdict(list)
            for d in json.load(open(tj)):
                name = d["kernel"].split("::")[-1].split("(")[0].replace("void ", "")
                name = name.split("<")[0] if name.startswith("prb_") else name
                seen[name].append(d)
            for name, ds in seen.items():
                n = len(ds)
                t[f"{cfg}:{name}"] = {"ncu_duration_us": round(sum(d["gpu__time_duration.sum"] for d in ds) / n, 2),
                                      "dram_traffic_bytes": int(sum(d["dram_traffic_bytes"] for d in ds) / n),
                                      "dram_read_bytes": int(sum(d["dram__bytes_read.sum"] for d in ds) / n),
                                      "dram_write_bytes": int(sum(d["dram__bytes_write.sum"] for d in ds) / n),
                                      "launches_captured": n,
                                      "source": f"profiles/{tag}.json (ncu --set full --clock-control none; cold caches, serialised)"}
                avg = lambda k: round(sum(d.get(k, 0) for d in ds) / n, 3)  # noqa: E731
                t[f"{cfg}:{name}"].update({
                    "l2_hit_rate_pct": avg("lts__t_sector_hit_rate.pct"),
                    "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                    "stalls_per_issue": {r: avg(f"smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio")
                                         for r in ("long_scoreboard", "short_scoreboard", "wait", "barrier", "membar")}})
            json.dump(t, open(tp, "w"), indent=1)
    if os.path.exists(os.path.join(g, "launches_s7.csv")):
        launches(os.path.join(g, "launches_s7.csv"), "r1_launch_list_bench_steps20_s7")
    if os.path.exists(os.path.join(g, "launches_r1b.csv")):
        launches(os.path.join(g, "launches_r1b.csv"), "r1_launch_list_bench_steps20_final")
    if os.path.exists(os.path.join(g, "launches_r1.csv")):
        launches(os.path.join(g, "launches_r1.csv"), "r1_launch_list_bench_steps20")
