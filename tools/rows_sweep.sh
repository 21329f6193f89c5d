#!/bin/bash
# batch-major row scans at C4: steps per lane x resident CTAs per SM (bench.py --only c4, by events after a flush)
for cfg in "8 4" "16 4" "16 6" "16 3" "8 3" "8 5"; do
  set -- $cfg
  python - $1 $2 <<'PY'
import sys, json, subprocess, os
spl, per = sys.argv[1], sys.argv[2]
code = f'''
import sys, json
sys.argv = ["bench.py", "--only", "c4", "--steps", "50", "--no-cpu"]
sys.path.insert(0, ".")
from mindrl_b200 import _ffi
_ffi.set_option("row_spl", {spl}); _ffi.set_option("row_ctas_per_sm", {per})
import runpy
runpy.run_path("bench.py", run_name="__main__")
'''
out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    v = d["variants"]
    print(f"spl {spl} ctas/sm {per}: gae_bm {v['gae_batch_major[B,T]']['ms']*1e3:.2f} us (b2b frac {v['gae_batch_major[B,T]']['back_to_back_frac']:.3f})  dr_bm {v['discounted_return_batch_major']['ms']*1e3:.2f} us (b2b {v['discounted_return_batch_major']['back_to_back_frac']:.3f})")
except Exception as e:
    print("failed", spl, per, e, out.stderr[-500:])
PY
done
