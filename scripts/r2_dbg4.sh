#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_dbg4.log
: > $LOG
for dbg in 0 8 9 10 11 12; do
  export COCLR_TMA_DBG=$dbg
  echo "---- DBG=$dbg" >> $LOG
  timeout 60 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 64 256 1 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 256 160 1 1 1 32 16 16 16 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py dgrad 256 128 1 1 1 32 16 16 16 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
done
cat $LOG
