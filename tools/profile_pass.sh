#!/bin/bash
# full validation + profiling pass (run under gpurun); outputs under gpurun_out/ with the given tag
TAG=${1:-s1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
python bench.py --steps 300 --warmup 10 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu > $O/bench_short_$TAG.json 2> $O/bench_short_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > $O/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
python tools/run_c2.py 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"prb_sample|prb_update_sub" -s 123 -c 4 \
    -o $O/prof_c2_$TAG -f python tools/run_c2.py 3 > $O/ncu_c2_$TAG.log 2>&1; echo "ncu c2 rc=$?"
python tools/run_c3.py 3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"prb_sample|prb_update_sub|gather_bulk" -s 123 -c 4 \
    -o $O/prof_c3_$TAG -f python tools/run_c3.py 3 > $O/ncu_c3_$TAG.log 2>&1; echo "ncu c3 rc=$?"
python tools/run_c4.py 1 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"scan_|norm_" -c 7 \
    -o $O/prof_c4_$TAG -f python tools/run_c4.py 1 > $O/ncu_c4_$TAG.log 2>&1; echo "ncu c4 rc=$?"
tail -2 $O/pytest_gpu_$TAG.log
