#!/bin/bash
mkdir -p gpurun_out
for grp in test_affine_split test_conv_forward test_conv_accumulate test_conv_dgrad test_conv_wgrad test_maxpool; do
  timeout 420 python -m pytest tests/test_conv_gpu.py -q -m gpu -k "$grp" --timeout 90 --timeout-method=thread \
      -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/pytest_$grp.log
  echo "== $grp exit ${PIPESTATUS[0]}"; tail -6 gpurun_out/pytest_$grp.log | cut -c1-300
done
