#!/bin/bash
# what the driver runs at round end, plus the ncu launch list of the bench command and the r50 bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider 2>&1 | tail -6
cp gpurun_out/test_diag.json gpurun_out/test_diag_full.json 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -3
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cat gpurun_out/bench_default.json
timeout 300 python bench.py --net r50 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r50.json 2> gpurun_out/bench_r50.err; echo "r50 exit $?"; head -c 330 gpurun_out/bench_r50.json; echo
COCLR_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.json 2>&1; echo "ncu exit $?"; wc -l gpurun_out/launches.csv
