#!/bin/bash
# CTA-pair (cta_group::2) mode of the TMA conv kernel: parity with short timeouts (a deadlock must not eat the box), then A/B
mkdir -p gpurun_out
LOG=gpurun_out/r2_pair.log
: > $LOG
export COCLR_TMA_PAIR=1
timeout 90 python -m pytest tests/test_conv_tma_gpu.py -x -q -p no:cacheprovider -k "pw_64_64_resident and forward" 2>&1 | tail -6 >> $LOG
echo "== pair smoke rc $?" >> $LOG
timeout 300 python -m pytest tests/test_conv_tma_gpu.py -q -p no:cacheprovider 2>&1 | tail -25 >> $LOG
echo "== pair tma suite rc $?" >> $LOG
timeout 300 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider -k "forward or dgrad or accumulate" 2>&1 | tail -8 >> $LOG
echo "== pair conv suite rc $?" >> $LOG
for pair in 1 0; do
  export COCLR_TMA_PAIR=$pair
  echo "---- PAIR=$pair" >> $LOG
  timeout 60 python tests/tools/run_one_conv.py s2d 3 64 1 4 4 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 60 python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 192 192 3 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 128 192 1 3 3 32 16 16 16 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 64 64 1 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 60 python tests/tools/run_one_conv.py fwd 256 160 1 1 1 32 16 16 16 >> $LOG 2>&1
done
cat $LOG
