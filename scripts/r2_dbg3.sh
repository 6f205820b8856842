#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_dbg3.log
: > $LOG
for dbg in 4 20; do
  export COCLR_TMA_DBG=$dbg
  echo "---- DBG=$dbg" >> $LOG
  timeout 60 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
done
cat $LOG
