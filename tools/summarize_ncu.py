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

def kernel_key(full_name):
    """short key of a kernel for the traffic table: template arguments kept for the copy / scan kernels"""
    name = full_name.split("::")[-1].split("(")[0].replace("void ", "").strip()
    if name.startswith("prb_sample_kernel") or name.startswith("prb_update"):
        name = name.split("<")[0]
    return name


def main(rnd, tag):
    """python tools/summarize_ncu.py r2 s1: gpurun_out/prof_{c2,c3,c4}_<tag>.ncu-rep + launches_<tag>.csv ->
    profiles/<rnd>_ncu_full_*_<tag>.json, profiles/<rnd>_launch_list_bench_steps20_<tag>.csv, profiles/<rnd>_traffic.json"""
    g = os.path.join(ROOT, "gpurun_out")
    traffic = {}
    tp = os.path.join(OUT, f"{rnd}_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    for cfg, rep, name in (("c2", f"prof_c2_{tag}.ncu-rep", f"{rnd}_ncu_full_c2_per_kernels_{tag}"),
                           ("c3", f"prof_c3_{tag}.ncu-rep", f"{rnd}_ncu_full_c3_per_kernels_{tag}"),
                           ("c4", f"prof_c4_{tag}.ncu-rep", f"{rnd}_ncu_full_c4_scan_kernels_{tag}")):
        path = os.path.join(g, rep)
        if not os.path.exists(path):
            print("missing", path)
            continue
        rows = report(path, name)
        seen = collections.defaultdict(list)
        for d in rows:
            print(name, d["kernel"][:70], d["gpu__time_duration.sum"], d.get("gpu__time_duration.sum_unit"),
                  round(d["dram_traffic_bytes"] / 1e6, 2), "MB")
            seen[kernel_key(d["kernel"])].append(d)
        for k, ds in seen.items():
            n = len(ds)
            avg = lambda key: round(sum(d.get(key, 0) for d in ds) / n, 3)  # noqa: E731
            traffic[f"{cfg}:{k}"] = {
                "ncu_duration_us": round(sum(d["gpu__time_duration.sum"] for d in ds) / n, 2),
                "dram_traffic_bytes": int(sum(d["dram_traffic_bytes"] for d in ds) / n),
                "dram_read_bytes": int(sum(d["dram__bytes_read.sum"] for d in ds) / n),
                "dram_write_bytes": int(sum(d["dram__bytes_write.sum"] for d in ds) / n),
                "launches_captured": n,
                "l2_hit_rate_pct": avg("lts__t_sector_hit_rate.pct"),
                "warps_active_pct": avg("sm__warps_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": avg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "stalls_per_issue": {r: avg(f"smsp__average_warps_issue_stalled_{r}_per_issue_active.ratio")
                                     for r in ("long_scoreboard", "short_scoreboard", "wait", "barrier", "membar")},
                "source": f"profiles/{name}.json (ncu --set full --clock-control none; cold caches, serialised)"}
    json.dump(traffic, open(tp, "w"), indent=1)
    lp = os.path.join(g, f"launches_{tag}.csv")
    if os.path.exists(lp):
        launches(lp, f"{rnd}_launch_list_bench_steps20_{tag}")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r2", sys.argv[2] if len(sys.argv) > 2 else "s1")
