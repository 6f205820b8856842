#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step17.log
: > $LOG
for sp in 12 24 49 98 148; do
  COCLR_WGRAD_SPLITS=$sp timeout 60 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 2>&1 | tail -1 | sed "s/^/splits=$sp /" >> $LOG
done
for sp in 16 37 74 148 256; do
  COCLR_WGRAD_SPLITS=$sp timeout 60 python tests/tools/run_one_conv.py wgrad 256 128 1 1 1 32 16 16 16 2>&1 | tail -1 | sed "s/^/splits=$sp /" >> $LOG
done
for sp in 4 8 16 32; do
  COCLR_WGRAD_SPLITS=$sp timeout 60 python tests/tools/run_one_conv.py wgrad 832 384 1 1 1 32 4 4 4 2>&1 | tail -1 | sed "s/^/splits=$sp /" >> $LOG
done
cat $LOG
