#!/bin/bash
# Runs the conv kernel parity tests in separate processes (a hung kernel only loses its own group).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in test_conv_forward test_conv_accumulate test_conv_dgrad test_conv_wgrad; do
  timeout 420 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "$grp" --timeout 90 --timeout-method=thread \
      -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/pytest_$grp.log
  echo "== $grp exit ${PIPESTATUS[0]}"; tail -15 gpurun_out/pytest_$grp.log
done
