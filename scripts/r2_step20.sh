#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step20.log
: > $LOG
timeout 200 python tests/tools/smoke_layerwise.py 2>&1 | grep -v Warn > gpurun_out/smoke_layers_default.txt
head -1 gpurun_out/smoke_layers_default.txt >> $LOG
grep "<<<" gpurun_out/smoke_layers_default.txt | head -12 >> $LOG
for v in "COCLR_TMA=0" "COCLR_TMA_NOSTACK=1" "COCLR_POOL333_REG=1 COCLR_POOL133_REG=1"; do
  echo "--- $v" >> $LOG
  env $v timeout 200 python tests/tools/smoke_layerwise.py 2>&1 | grep -v Warn | head -1 >> $LOG
done
cat $LOG
