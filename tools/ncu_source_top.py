"""top stall instructions of one kernel in an .ncu-rep:  ncu_source_top.py <report> <launch index> [n]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
h = rows[1]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
data = [r for r in rows[2:] if len(r) == len(h) and (r[si] or "0").isdigit() and r[0].startswith("0x")]
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {h[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
    print("  ", k, v, round(100 * v / tot, 1), "%")
idx = sorted(range(len(data)), key=lambda j: -int(data[j][si] or 0))[:n]
for j in sorted(idx):
    r = data[j]
    reasons = {h[i][6:]: int(r[i] or 0) for i in stall_cols if int(r[i] or 0) > 0}
    top = sorted(reasons.items(), key=lambda kv: -kv[1])[:3]
    print(f"{j:5d} {r[src].strip()[:64]:64s} smp {int(r[si] or 0):5d} ({100 * int(r[si] or 0) / tot:4.1f}%) exec {r[ie]:>7s} {top}")
