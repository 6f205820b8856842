#!/bin/bash
mkdir -p gpurun_out
PREC=${1:-parity}
timeout 600 python bench.py --steps 4 --warmup 3 --breakdown --no-cpu-baseline --no-e2e --precision $PREC > gpurun_out/bench_$PREC.json 2> gpurun_out/bench_$PREC.err
echo "== bench $PREC exit $?"; cut -c1-200 gpurun_out/bench_$PREC.json; grep -v Warning gpurun_out/bench_$PREC.err | head -80
