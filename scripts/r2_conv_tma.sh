#!/bin/bash
# round 2: parity of the TMA-staged conv kernel, then A/B timings against the cp.async gather kernel
mkdir -p gpurun_out
LOG=gpurun_out/r2_conv_tma.log
: > $LOG
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I coclr_b200/csrc -o /tmp/umma_shift_probe tests/tools/probes/umma_shift_probe.cu >> $LOG 2>&1 && timeout 120 /tmp/umma_shift_probe > gpurun_out/r2_umma_shift_probe.txt 2>&1
echo "== probe rc $?" >> $LOG
COCLR_TMA_DEBUG=1 timeout 400 python -m pytest tests/test_conv_tma_gpu.py -q -p no:cacheprovider -k "forward" 2>&1 | tail -40 >> $LOG
echo "== tma forward rc $?" >> $LOG
COCLR_TMA_DEBUG=1 timeout 400 python -m pytest tests/test_conv_tma_gpu.py -q -p no:cacheprovider -k "dgrad or bitwise" 2>&1 | tail -40 >> $LOG
echo "== tma dgrad rc $?" >> $LOG
timeout 900 python -m pytest tests/test_conv_gpu.py -q -p no:cacheprovider 2>&1 | tail -15 >> $LOG
echo "== conv suite rc $?" >> $LOG
for tma in 1 0; do
  export COCLR_TMA=$tma
  echo "---- COCLR_TMA=$tma" >> $LOG
  timeout 120 python tests/tools/run_one_conv.py fwd 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 192 192 3 1 1 32 16 32 32 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py fwd 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  STRIDE=2,1,1 timeout 120 python tests/tools/run_one_conv.py dgrad 64 64 7 1 1 32 32 64 64 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py dgrad 64 192 1 3 3 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py dgrad 192 192 3 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 64 64 1 1 1 32 16 32 32 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 256 160 1 1 1 32 16 16 16 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py dgrad 256 128 1 1 1 32 16 16 16 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 128 192 1 3 3 32 16 16 16 >> $LOG 2>&1
  timeout 120 python tests/tools/run_one_conv.py fwd 480 192 1 1 1 32 8 8 8 >> $LOG 2>&1
done
tail -60 $LOG
