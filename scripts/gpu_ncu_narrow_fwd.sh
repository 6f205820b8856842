#!/bin/bash
mkdir -p gpurun_out
STRIDE=2,1,1 ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 3 -c 1 -o gpurun_out/ncu_r01_fwd_tm7s2_v2 -f python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 3 2 > /dev/null 2>&1
ls -la gpurun_out/ncu_r01_fwd_tm7s2_v2.ncu-rep
