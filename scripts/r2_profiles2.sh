#!/bin/bash
# round-2 evidence for the TMA-staged weight-gradient kernel: --set full captures on its two heaviest shapes and the
# launch list of the bench command with the final code
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
WS=1 timeout 300 $N -k regex:wgrad_tma -s 3 -c 1 -o gpurun_out/ncu_r02_wgrad_tma_conv2c1 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
WS=1 STRIDE=2,1,1 timeout 300 $N -k regex:wgrad_tma -s 3 -c 1 -o gpurun_out/ncu_r02_wgrad_tma_stem2 python tests/tools/run_one_conv.py wgrad 64 64 7 1 1 32 32 64 64 3 2 > /dev/null 2>&1
WS=1 timeout 300 $N -k regex:wgrad_reduce -s 3 -c 1 -o gpurun_out/ncu_r02_wgrad_reduce_conv2c1 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
COCLR_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-gpu --no-parity --no-mixed > gpurun_out/r02b_bench_under_ncu.json 2>&1
ls -la gpurun_out/ncu_r02_wgrad*.ncu-rep; wc -l gpurun_out/r02b_launches.csv
WS=1 STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py wgrad 64 64 7 1 1 32 32 64 64 2>&1 | tail -1
WS=1 timeout 60 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 2>&1 | tail -1
