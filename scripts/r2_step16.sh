#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r2_step16.log
: > $LOG
COCLR_TMA_DEBUG=1 timeout 90 python -m pytest tests/test_wgrad_tma_gpu.py -q -p no:cacheprovider -x -k "pw_64_64 and 3 and fp16" 2>&1 | tail -15 >> $LOG
echo "== first case rc ${PIPESTATUS[0]}" >> $LOG
if grep -q "passed" $LOG && ! grep -q "failed" $LOG; then
  timeout 400 python -m pytest tests/test_wgrad_tma_gpu.py -q -p no:cacheprovider 2>&1 | tail -30 >> $LOG
  echo "== all cases rc ${PIPESTATUS[0]}" >> $LOG
  timeout 300 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k wgrad 2>&1 | tail -3 >> $LOG
  for m in 1 0; do
    echo "--- COCLR_WGRAD_TMA=$m" >> $LOG
    COCLR_WGRAD_TMA=$m SPLITS=49 timeout 60 python tests/tools/run_one_conv.py wgrad 64 192 1 3 3 32 16 32 32 2>&1 | tail -1 >> $LOG
    COCLR_WGRAD_TMA=$m SPLITS=49 timeout 60 python tests/tools/run_one_conv.py wgrad 192 192 3 1 1 32 16 32 32 2>&1 | tail -1 >> $LOG
    COCLR_WGRAD_TMA=$m SPLITS=49 timeout 60 python tests/tools/run_one_conv.py wgrad 256 128 1 1 1 32 16 16 16 2>&1 | tail -1 >> $LOG
    COCLR_WGRAD_TMA=$m SPLITS=24 timeout 60 python tests/tools/run_one_conv.py wgrad 112 224 1 3 3 32 8 8 8 2>&1 | tail -1 >> $LOG
    COCLR_WGRAD_TMA=$m SPLITS=24 timeout 60 python tests/tools/run_one_conv.py wgrad 832 384 1 1 1 32 4 4 4 2>&1 | tail -1 >> $LOG
  done
fi
cat $LOG
