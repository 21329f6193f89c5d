#!/bin/bash
# N = 1 and N = <max_n> only (C2 at 20 and 200 steps, C5 at 200): the end points of the scaling curve on one box
tag=${1:-r2}
maxn=${2:-8}
mkdir -p gpurun_out
port=29700
for n in 1 $maxn; do
  for what in "c2 20" "c2 200" "c5 200"; do
    set -- $what
    port=$((port+1))
    out=gpurun_out/${tag}_scale_n${n}_$1_s$2
    if [ $n -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps $2 --warmup 5 --only $1 --no-cpu > $out.json 2> $out.err
    else
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps $2 --warmup 5 --only $1 --no-cpu > $out.json 2> $out.err
    fi
    python - $out.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    ex = d.get("exchange") or {}
    fr = d.get("free_running") or {}
    print(sys.argv[1], "N", d["n_gpus"], "value %.2fM" % (d["value"] / 1e6), "us/step %.2f" % (d["ms_per_step"] * 1e3),
          "e2e %.2fM" % (d["e2e"]["value"] / 1e6), "parity", (d.get("parity_check") or {}).get("ok"),
          "wait_us %.2f waited %s" % (ex.get("exchange_wait_us", 0), ex.get("calls_that_waited")),
          "free %.2fM wait_us %.2f" % ((fr.get("value") or 0) / 1e6, fr.get("exchange_wait_us") or 0),
          "c5eff", (d.get("c5") or {}).get("weak_scaling_efficiency_vs_same_run_single_gpu"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
