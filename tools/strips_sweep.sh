#!/bin/bash
# time-major strip scans at C4: strip width (bench.py --only c4, by events after a flush and back to back)
for cols in 32 28; do
  python - $cols <<'PY'
import sys, json, subprocess
cols = sys.argv[1]
code = f'''
import sys
sys.argv = ["bench.py", "--only", "c4", "--steps", "50", "--no-cpu"]
sys.path.insert(0, ".")
from mindrl_b200 import _ffi
_ffi.set_option("strip_cols", {cols})
import runpy
runpy.run_path("bench.py", run_name="__main__")
'''
out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    v = d["variants"]
    for k in ("gae_time_major[T,B]", "discounted_return_time_major", "gae+normalize_fused_moments[T,B]"):
        print(f"cols {cols}: {k}: {v[k]['ms']*1e3:.2f} us frac {v[k]['frac']:.3f} b2b {v[k].get('back_to_back_frac', 0):.3f}")
    print("   scan_error max_abs_over_norm", d["scan_error"]["gae_time_major_adv"]["max_abs_over_norm"])
except Exception as e:
    print("failed", cols, e, out.stderr[-800:])
PY
done
