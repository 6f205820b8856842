#!/bin/bash
# ncu --set full captures of the dominant kernel on its three heaviest launch shapes of the step
mkdir -p gpurun_out
python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32
python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32
python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 16 64 64
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/ncu_r01_fwd_conv2c -f python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 3 -c 1 -o gpurun_out/ncu_r01_wgrad_conv2c -f python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/ncu_r01_fwd_tm7 -f python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 16 64 64 3 2 > /dev/null 2>&1
ls -la gpurun_out/ncu_r01_*.ncu-rep
