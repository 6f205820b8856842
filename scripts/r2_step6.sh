#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step6.log
: > $LOG
timeout 600 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "bn_relu_backward or maxpool" 2>&1 | tail -8 >> $LOG
echo "== unit rc $?" >> $LOG
for old in "" 1; do
  if [ -n "$old" ]; then export COCLR_POOL333_OLD=1; fi
  echo "---- old=$old" >> $LOG
  timeout 120 python tests/tools/run_one_pool.py 256 16 16 16 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 192 16 16 16 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 480 8 8 8 3 1 1 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_pool.py 832 4 4 4 3 1 1 >> $LOG 2>&1
done
unset COCLR_POOL333_OLD
timeout 120 python tests/tools/run_one_pool.py 64 16 64 64 1 2 1 >> $LOG 2>&1
timeout 120 python tests/tools/run_one_pool.py 192 16 32 32 1 2 1 >> $LOG 2>&1
timeout 120 python tests/tools/run_one_pool.py 480 16 16 16 3 2 1 >> $LOG 2>&1
cat $LOG
