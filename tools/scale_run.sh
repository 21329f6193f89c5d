#!/bin/bash
# Scaling run on one multi-GPU box: sharded PER (C2, C5) at N = 1, 2, 4, 8 back to back, the world-N parity test,
# and the exchange probe.  Usage: tools/scale_run.sh <tag> [max_n]   (outputs under gpurun_out/)
tag=${1:-r2}
maxn=${2:-8}
mkdir -p gpurun_out
port=29600
MRL_TEST_WORLD=$maxn MRL_TEST_LOG=gpurun_out/${tag}_shard_world${maxn}.log timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=900 2>&1 | tail -5 > gpurun_out/${tag}_gputest_multi${maxn}.log
tail -2 gpurun_out/${tag}_gputest_multi${maxn}.log
grep "SHARD_" gpurun_out/${tag}_shard_world${maxn}.log
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
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
    print(sys.argv[1], "N", d["n_gpus"], "value %.2fM" % (d["value"] / 1e6), "us/step %.2f" % (d["ms_per_step"] * 1e3),
          "e2e %.2fM" % (d["e2e"]["value"] / 1e6), "parity", (d.get("parity_check") or {}).get("ok"),
          "wait_us %.2f waited %s" % (ex.get("exchange_wait_us", 0), ex.get("calls_that_waited")),
          "free %.2fM" % (((d.get("free_running") or {}).get("value") or 0) / 1e6),
          "c5eff", (d.get("c5") or {}).get("weak_scaling_efficiency_vs_same_run_single_gpu"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
