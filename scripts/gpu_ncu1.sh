#!/bin/bash
mkdir -p gpurun_out
# (a) epilogue-bound dgrad of a 1x1 conv (Cout=16 -> Cin=192), (b) the big 3x3 forward, (c) its wgrad
python tests/tools/run_one_conv.py dgrad 192 16 1 1 1 32 16 16 16
python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32
python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/ncu_dgrad_small -f python tests/tools/run_one_conv.py dgrad 192 16 1 1 1 32 16 16 16 3 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/ncu_fwd_2c -f python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_wgrad -s 3 -c 1 -o gpurun_out/ncu_wgrad_2c -f python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 3 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
