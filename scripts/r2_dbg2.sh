#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_dbg2.log
: > $LOG
for v in 0 1 2; do
 for dbg in 0 8 12; do
  export COCLR_TMA_DBG=$dbg COCLR_TMA_VARIANT=$v
  echo "---- VARIANT=$v DBG=$dbg" >> $LOG
  timeout 60 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 192 192 3 1 1 32 16 32 32 >> $LOG 2>&1
 done
done
unset COCLR_TMA_VARIANT
export COCLR_TMA_DBG=12
timeout 60 python tests/tools/run_one_conv.py fwd 64 64 1 1 1 32 16 32 32 >> $LOG 2>&1
timeout 60 python tests/tools/run_one_conv.py fwd 256 160 1 1 1 32 16 16 16 >> $LOG 2>&1
grep -v "^Traceback\|^  File\|^    " $LOG
