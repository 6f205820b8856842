#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_pool.log
: > $LOG
timeout 300 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "maxpool" 2>&1 | tail -5 >> $LOG
echo "== pool tests rc $?" >> $LOG
for reg in "" 1; do
  if [ -n "$reg" ]; then export COCLR_POOL333_REG=1; fi
  echo "---- register-only kernel=$reg" >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 256 16 16 16 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 192 16 16 16 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 480 8 8 8 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 528 8 8 8 3 1 1 >> $LOG 2>&1
done
cat $LOG
