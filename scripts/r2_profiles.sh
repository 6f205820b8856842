#!/bin/bash
# round-2 evidence: ncu launch list of the bench command, --set full captures of the dominant kernel on its heaviest
# shapes and of the HBM-bound kernels the north star names (EMA, enqueue, NCE, Adam, BN passes)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on -f"
STRIDE=2,1,1 timeout 300 $N -k regex:conv_tma -s 3 -c 1 -o gpurun_out/ncu_r02_tma_stem2_fwd python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 3 2 > /dev/null 2>&1
STRIDE=2,1,1 timeout 300 $N -k regex:conv_tma -s 3 -c 1 -o gpurun_out/ncu_r02_tma_stem2_dgrad python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64 3 2 > /dev/null 2>&1
timeout 300 $N -k regex:conv_tma -s 3 -c 1 -o gpurun_out/ncu_r02_tma_conv2c1_fwd python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
timeout 300 $N -k regex:conv_tma -s 3 -c 1 -o gpurun_out/ncu_r02_tma_conv2c2_fwd python tests/tools/run_one_conv.py fwd 192 192 3 1 1 32 16 32 32 3 2 > /dev/null 2>&1
timeout 300 $N -k regex:conv_tma -s 3 -c 1 -o gpurun_out/ncu_r02_tma_s2d python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 3 2 > /dev/null 2>&1
timeout 300 $N -k regex:conv_wgrad -s 3 -c 1 -o gpurun_out/ncu_r02_wgrad_conv2c1 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
# HBM-bound kernels inside a real step (graphs off so that ncu sees the launches); second step of the run
COCLR_GRAPHS=0 timeout 900 $N -k regex:"ema_kernel|enqueue_kernel|nce_fwd_kernel|nce_bwd_kernel|adam_kernel|pack_weights_batch|pack_input_s2d" -s 7 -c 7 -o gpurun_out/ncu_r02_step_small python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-gpu --no-parity --no-mixed > /dev/null 2>&1
COCLR_GRAPHS=0 timeout 900 $N -k regex:"bn_apply_split|bn_bwd_apply|bn_bwd_reduce" -s 180 -c 3 -o gpurun_out/ncu_r02_step_bn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-gpu --no-parity --no-mixed > /dev/null 2>&1
COCLR_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-stock-gpu --no-parity --no-mixed > gpurun_out/r02_bench_under_ncu.json 2>&1
ls -la gpurun_out/ncu_r02_*.ncu-rep; wc -l gpurun_out/r02_launches.csv
